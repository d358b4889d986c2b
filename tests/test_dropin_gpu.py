"""GPU: the drop-in boundary is real.  libxaac_b200/dropin/_build/xaacdec_b200 (make dropin) is the reference's OWN testbench
and decoder library, unmodified, linked with the link-time stage overrides of libxaac_b200/dropin/ixheaacd_b200_glue.c
(ld --wrap=ixheaacd_imdct_process / ixheaacd_sbr_dec / ixheaacd_fd_frm_dec / ixheaacd_channel_pair_process) against libxaac_b200.so: the reference's bitstream
parser calls the B200 kernels.  Whole files of >= 3000 frames made by the reference encoder are decoded by it and by the plain
reference decoder (oracle/_ref/xaacdec); the WAV files must be byte-identical and, for the encoder's default settings, not one
stage call may fall back to the reference's own code."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
B200 = os.path.join(ROOT, "libxaac_b200", "dropin", "_build", "xaacdec_b200")


def _need():
    for p in (B200, os.path.join(REFDIR, "xaacdec"), os.path.join(REFDIR, "xaacenc")):
        if not os.path.exists(p):
            pytest.skip(f"{os.path.relpath(p, ROOT)} not built (make ref && make dropin; needs /root/reference at build time)")


def _synth_wav(path, fs, seconds, ch, seed):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden as mg
    mg.write_wav(path, mg.synth(fs, seconds, ch, seed), fs)


def _run(cmd, env=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert r.returncode == 0, (cmd, r.stdout.decode(errors="replace")[-1500:], r.stderr.decode(errors="replace")[-1500:])
    return r.stderr.decode(errors="replace")


CASES = [
    # name, encoder args, sample rate, channels, seconds (>= 3000 frames), decoder args, expected GPU stage counters
    ("aac_lc_stereo", ["-aot:2", "-adts:1", "-br:128000"], 44100, 2, 71.0, [], dict(imdct=6000)),
    ("heaac_v1_mono", ["-aot:5", "-adts:1", "-br:32000"], 44100, 1, 141.0, ["-esbr:0"], dict(imdct=3000, hq=3000)),
    ("heaac_v1_stereo", ["-aot:5", "-adts:1", "-br:48000"], 48000, 2, 130.0, ["-esbr:0"], dict(imdct=6000, lp=6000)),
    ("heaac_v2", ["-aot:29", "-adts:1", "-br:32000"], 44100, 2, 141.0, ["-esbr:0"], dict(imdct=3000, ps=3000)),
]


@pytest.mark.parametrize("name,enc,fs,ch,secs,dec,want", CASES, ids=[c[0] for c in CASES])
def test_whole_file_through_reference_parser(tmp_path, name, enc, fs, ch, secs, dec, want):
    _need()
    wav = str(tmp_path / "in.wav")
    _synth_wav(wav, fs, secs, ch, 100 + len(name))
    bits = str(tmp_path / (name + ".aac"))
    _run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{bits}"] + enc)
    ref_wav, our_wav = str(tmp_path / "ref.wav"), str(tmp_path / "b200.wav")
    _run([os.path.join(REFDIR, "xaacdec"), f"-ifile:{bits}", f"-ofile:{ref_wav}"] + dec)
    log = _run([B200, f"-ifile:{bits}", f"-ofile:{our_wav}"] + dec, env=dict(os.environ, IXHEAACD_B200_STATS="1"))
    m = re.search(r"imdct_process: (\d+) on the GPU, (\d+) by the reference; sbr_dec: (\d+) HQ \+ (\d+) HQ/PS \+ (\d+) LP on the GPU, "
                  r"(\d+) by the reference; fd_frm_dec: (\d+) on the GPU, (\d+) by the reference", log)
    assert m, log[-800:]
    imdct, imdct_ref, hq, ps, lp, sbr_ref, fd, fd_ref = map(int, m.groups())
    a, b = open(ref_wav, "rb").read(), open(our_wav, "rb").read()
    assert len(a) == len(b) and len(a) > 44 + 2 * ch * 1024 * 2900
    if a != b:
        x, y = np.frombuffer(a[44:], np.int16), np.frombuffer(b[44:], np.int16)
        bad = np.flatnonzero(x != y)
        raise AssertionError(f"{name}: {bad.size} of {x.size} samples differ, first at sample {bad[0]} (frame {bad[0] // (ch * 1024)}), "
                             f"max |diff| {np.abs(x.astype(int) - y.astype(int)).max()}")
    assert imdct_ref == 0 and sbr_ref == 0 and fd_ref == 0, f"{name}: stage calls fell back to the reference: {m.group(0)}"
    assert imdct >= want.get("imdct", 0) and hq >= want.get("hq", 0) and ps >= want.get("ps", 0) and lp >= want.get("lp", 0), m.group(0)
    if "hq" not in want:
        assert hq == 0
    if "lp" not in want:
        assert lp == 0
    if "ps" not in want:
        assert ps == 0
    # the pre-IMDCT spectral stage (ixheaacd_channel_pair_process: M/S, intensity, TNS) of every element ran on the GPU as well
    m2 = re.search(r"channel_pair_process: (\d+) on the GPU \((\d+) with M/S or intensity bands, (\d+) with TNS, \d+ with PNS\), (\d+) by the reference", log)
    assert m2, log[-600:]
    cpp, cpp_ms, cpp_tns, cpp_ref = map(int, m2.groups())
    assert cpp >= 3000 and cpp_ref == 0 and cpp_tns >= 50, m2.group(0)
    if name in ("aac_lc_stereo", "heaac_v1_stereo"):
        assert cpp_ms >= 2000, m2.group(0)
    # ... and so did the SBR side-info dequantisation (ixheaacd_dec_sbrdata) of every SBR frame
    m3 = re.search(r"dec_sbrdata: (\d+) on the GPU \((\d+) coupled pairs, (\d+) with a concealed channel\), (\d+) by the reference", log)
    assert m3, log[-600:]
    sd, sd_coupled, sd_conc, sd_ref = map(int, m3.groups())
    if name != "aac_lc_stereo":
        assert sd >= 2900 and sd_ref == 0, m3.group(0)
    m4 = re.search(r"decode_ps_data: (\d+) on the GPU", log)
    assert m4, log[-600:]
    if name == "heaac_v2":
        assert int(m4.group(1)) >= 2900, m4.group(0)


@pytest.mark.parametrize("name,extra", [("usac", []), ("usac_hbe", ["-harmonic_sbr:1"])])
def test_usac_through_reference_parser(tmp_path, name, extra):
    """xHE-AAC (USAC, aot 42, ccfl 1024, stereo 32 kHz): the FD core transform of every frame runs on the GPU behind
    ixheaacd_fd_frm_dec and the float eSBR branch behind ixheaacd_sbr_dec — xaac_b200_esbr_dec_dev, or xaac_b200_esbr_dec_hbe_dev
    when the stream carries the harmonic transposer.  On reset frames and on frames where sbr_patching_mode changes the glue runs the
    stage in its two halves and rebuilds the limiter tables (ixheaacd_createlimiterbands) in between, so they stay on the GPU too.  The float path is graded at +-1 LSB
    (SURVEY 8c); the decode is expected to be byte-identical."""
    _need()
    wav = str(tmp_path / "in.wav")
    _synth_wav(wav, 32000, 100.0, 2, 77)
    bits = str(tmp_path / "usac.mp4")
    _run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{bits}", "-aot:42", "-br:64000", "-ccfl_idx:3"] + extra)
    meta = str(tmp_path / "usac.txt")
    ref_wav, our_wav = str(tmp_path / "ref.wav"), str(tmp_path / "b200.wav")
    _run([os.path.join(REFDIR, "xaacdec"), f"-ifile:{bits}", f"-ofile:{ref_wav}", f"-imeta:{meta}", "-mp4:1"])
    log = _run([B200, f"-ifile:{bits}", f"-ofile:{our_wav}", f"-imeta:{meta}", "-mp4:1"], env=dict(os.environ, IXHEAACD_B200_STATS="1"))
    m = re.search(r"fd_frm_dec: (\d+) on the GPU, (\d+) by the reference; eSBR sbr_dec: (\d+) \+ (\d+) with HBE \+ \d+ with PS on the GPU[^,]*, (\d+) by the reference", log)
    assert m, log[-800:]
    fd, fd_ref, es, es_hbe, es_ref = map(int, m.groups())
    a, b = open(ref_wav, "rb").read(), open(our_wav, "rb").read()
    assert len(a) == len(b) and len(a) > 44 + 4 * 1024 * 1500
    x, y = np.frombuffer(a[44:], np.int16).astype(np.int32), np.frombuffer(b[44:], np.int16).astype(np.int32)
    bad = np.flatnonzero(x != y)
    assert bad.size == 0 or np.abs(x - y).max() <= 1, (f"{name}: {bad.size} samples differ, first at {bad[0]} (frame {bad[0] // 4096}), "
                                                       f"max |diff| {np.abs(x - y).max()}")
    assert bad.size == 0, f"{name}: within +-1 LSB but not identical: {bad.size} samples"
    assert fd >= 2 * 1500 and fd_ref == 0, m.group(0)
    # the one frame left to the reference is the stream's first, where sbr_mode is still UNKNOWN_SBR and ixheaacd_sbr_env_calc skips its
    # per-envelope estimates (it then works on whatever its scratch buffer holds)
    if extra == ["-harmonic_sbr:1"]:
        assert es_hbe >= 2 * 1500 and es_ref <= 1, m.group(0)
    else:
        assert es >= 2 * 1500 and es_hbe == 0 and es_ref <= 1, m.group(0)


@pytest.mark.parametrize("name,enc,fs,ch,secs", [("heaac_v1_stereo_default_mode", ["-aot:5", "-adts:1", "-br:48000"], 48000, 2, 70.0),
                                                 ("heaac_v1_mono_default_mode", ["-aot:5", "-adts:1", "-br:32000"], 44100, 1, 70.0)])
def test_legacy_heaac_in_default_esbr_mode(tmp_path, name, enc, fs, ch, secs):
    """HE-AACv1 decoded with the reference's DEFAULT flags (no -esbr:0): the float eSBR branch with the harmonic transposer
    forced on (decoder/ixheaacd_sbrdecoder.c:400-403) — IMDCT and the eSBR + HBE stage on the GPU behind the reference parser, the
    PCM16 <-> float hand-overs stay the reference's (ixheaacd_api.c:3384-3437, ixheaacd_samples_sat)."""
    _need()
    wav = str(tmp_path / "in.wav")
    _synth_wav(wav, fs, secs, ch, 5 + len(name))
    bits = str(tmp_path / (name + ".aac"))
    _run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{bits}"] + enc)
    ref_wav, our_wav = str(tmp_path / "ref.wav"), str(tmp_path / "b200.wav")
    _run([os.path.join(REFDIR, "xaacdec"), f"-ifile:{bits}", f"-ofile:{ref_wav}"])
    log = _run([B200, f"-ifile:{bits}", f"-ofile:{our_wav}"], env=dict(os.environ, IXHEAACD_B200_STATS="1"))
    m = re.search(r"imdct_process: (\d+) on the GPU, (\d+) by the reference.*eSBR sbr_dec: (\d+) \+ (\d+) with HBE \+ \d+ with PS on the GPU[^,]*, (\d+) by the reference", log)
    assert m, log[-800:]
    imdct, imdct_ref, es, es_hbe, es_ref = map(int, m.groups())
    a, b = open(ref_wav, "rb").read(), open(our_wav, "rb").read()
    assert len(a) == len(b) and len(a) > 100000
    x, y = np.frombuffer(a[44:], np.int16).astype(np.int32), np.frombuffer(b[44:], np.int16).astype(np.int32)
    bad = np.flatnonzero(x != y)
    assert bad.size == 0, f"{name}: {bad.size} of {x.size} samples differ, first at {bad[0]}, max |diff| {np.abs(x - y).max()}; {m.group(0)}"
    assert imdct_ref == 0 and es_hbe >= ch * 1400 and es_ref == 0, m.group(0)  # not one stage call falls back


def test_heaac_v2_in_default_esbr_mode(tmp_path):
    """HE-AACv2 (mono core + SBR + PS) decoded with the reference's DEFAULT flags: the float eSBR branch with the harmonic
    transposer forced on AND the float parametric stereo (ixheaacd_esbr_apply_ps) — analysis bank, transposer, HF generator,
    envelope adjuster, PS kernel and both synthesis banks on the GPU behind the reference parser; the mixing matrices are
    prepared by the glue with the C library the reference itself calls."""
    _need()
    name, fs, secs = "heaac_v2_default_mode", 44100, 70.0
    wav = str(tmp_path / "in.wav")
    _synth_wav(wav, fs, secs, 2, 31)
    bits = str(tmp_path / (name + ".aac"))
    _run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{bits}", "-aot:29", "-adts:1", "-br:32000"])
    ref_wav, our_wav = str(tmp_path / "ref.wav"), str(tmp_path / "b200.wav")
    _run([os.path.join(REFDIR, "xaacdec"), f"-ifile:{bits}", f"-ofile:{ref_wav}"])
    log = _run([B200, f"-ifile:{bits}", f"-ofile:{our_wav}"], env=dict(os.environ, IXHEAACD_B200_STATS="1"))
    m = re.search(r"imdct_process: (\d+) on the GPU, (\d+) by the reference.*eSBR sbr_dec: (\d+) \+ (\d+) with HBE \+ (\d+) with PS on the GPU[^,]*, "
                  r"(\d+) by the reference", log)
    assert m, log[-800:]
    imdct, imdct_ref, es, es_hbe, es_ps, es_ref = map(int, m.groups())
    a, b = open(ref_wav, "rb").read(), open(our_wav, "rb").read()
    assert len(a) == len(b) and len(a) > 100000
    x, y = np.frombuffer(a[44:], np.int16).astype(np.int32), np.frombuffer(b[44:], np.int16).astype(np.int32)
    bad = np.flatnonzero(x != y)
    assert bad.size == 0, f"{name}: {bad.size} of {x.size} samples differ, first at {bad[0]}, max |diff| {np.abs(x - y).max()}; {m.group(0)}"
    assert imdct_ref == 0 and es_ps >= 1400 and es_ref == 0, m.group(0)  # not one stage call falls back
    assert np.abs(x[0::2] - x[1::2]).max() > 100  # a real stereo image came out of the mono core
