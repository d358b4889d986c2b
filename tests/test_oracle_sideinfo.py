"""CPU check that pins the side-info golden records (tests/golden/sbr_sideinfo.npz) on the compiled reference: the committed
outputs are what ixheaacd_dec_sbrdata / ixheaacd_decode_ps_data (oracle/_ref, through oracle/ref_shim_sd.c) produce from the
committed inputs, and the record generators reach the branches the GPU tests rely on."""
import os

import numpy as np

from tests import oracle_util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sbr_sideinfo.npz")


def test_golden_records_match_compiled_reference(ref):
    g = np.load(GOLDEN)
    assert np.array_equal(ref.dec_sbrdata_batch(g["records_in"]), g["records_out"])
    assert np.array_equal(ref.decode_ps_data_batch(g["ps_records_in"]), g["ps_records_out"])


def test_golden_inputs_are_what_the_generators_produce():
    g = np.load(GOLDEN)
    assert np.array_equal(oracle_util.synth_sbrdata_records(400, 2024), g["records_in"])
    assert np.array_equal(oracle_util.synth_psdata_records(400, 2025), g["ps_records_in"])


def test_generator_reaches_every_branch(ref):
    rec = oracle_util.synth_sbrdata_records(3000, 5)
    out = ref.dec_sbrdata_batch(rec)
    c0 = oracle_util.SD_CH
    S = oracle_util.SDC
    ok = out[:, 2] == 0
    assert ok.sum() > 2500 and (out[:, 2] == 1).sum() > 20                      # fatal timing errors are reported
    concealed = (rec[:, c0 + S["ERR_FLAG"]] == 0) & (out[:, c0 + S["ERR_FLAG"]] != 0) & ok
    assert concealed.sum() > 200                                                # concealment incl. range-check retries
    assert ((out[:, c0 + S["ERR_FLAG_PREV"]] != 0) & (rec[:, c0 + S["ERR_FLAG_PREV"]] == 0)).sum() > 100  # timing compensation
    coupled = (rec[:, 0] == 2) & (rec[:, c0 + S["COUPLING"]] == 1) & ok
    assert coupled.sum() > 300 and (rec[:, 1] == 1).sum() > 300                 # coupled pairs, shared headers
    ps = oracle_util.synth_psdata_records(2000, 6)
    pout = ref.decode_ps_data_batch(ps)
    P = oracle_util.PSD
    assert (pout[:, P["NUM_ENV"]] > ps[:, P["NUM_ENV"]]).sum() > 100            # "no data" frames and extended variable borders
    assert (ps[:, P["IID_MODE"]] == 2).sum() > 300 and (pout[:, P["DATA_PRESENT"]] == 0).all()
