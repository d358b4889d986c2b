"""Golden records for the SBR / PS side-info dequantisation: seeded XAAC_SD_* and XAAC_PSD_* records run through the compiled
reference (oracle/_ref/libxaac_ref.so : ref_dec_sbrdata_batch -> ixheaacd_dec_sbrdata, ref_decode_ps_data_batch ->
ixheaacd_decode_ps_data).  Run in the build container (needs `make ref`);
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
ps_rec = oracle_util.synth_psdata_records(400, 2025)
ps_out = ref.decode_ps_data_batch(ps_rec)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sbr_sideinfo.npz"), records_in=rec, records_out=out, ps_records_in=ps_rec,
                    ps_records_out=ps_out)
print("wrote", rec.shape, "error codes", np.unique(out[:, 2], return_counts=True))
