"""Golden records for the SBR side-info dequantisation: seeded XAAC_SD_* records run through the compiled reference
(oracle/_ref/libxaac_ref.so : ref_dec_sbrdata_batch -> ixheaacd_dec_sbrdata).  Run in the build container (needs `make ref`);
writes tests/golden/sbr_sideinfo.npz, which the GPU box checks against without /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_util  # noqa: E402

ref = oracle_util.Ref.try_load()
assert ref is not None, "build oracle/_ref first (make ref)"
rec = oracle_util.synth_sbrdata_records(400, 2024)
out = ref.dec_sbrdata_batch(rec)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sbr_sideinfo.npz"), records_in=rec, records_out=out)
print("wrote", rec.shape, "error codes", np.unique(out[:, 2], return_counts=True))
