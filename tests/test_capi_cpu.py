"""CPU suite, part 3: the C-ABI library loads, exports every symbol include/avsim.h declares, its GPU-free
helpers agree with the oracle, and it fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from auvsl_neural_ode_b200 import capi, synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "avsim.h")).read()
    declared = set(re.findall(r"\b(avs_[a-z0-9_]+)\s*\(", hdr)) - {"avs_ctx"}
    assert declared, "no declarations parsed"
    lib = C.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/avsim.h but not exported by libavsim.so"
    assert declared == set(capi.EXPORTS), (declared ^ set(capi.EXPORTS))


def test_static_helpers_match_oracle():
    assert np.array_equal(capi.nn_default_params(), po.nn_default_params())
    assert np.array_equal(capi.bekker_default_params(), po.bekker_default_params())
    np.testing.assert_allclose(capi.whole_body_com(), po.whole_body_com(), rtol=1e-15)
    rng = np.random.default_rng(0)
    s = rng.normal(size=21)
    np.testing.assert_allclose(capi.init_state_com(s), po.init_state_com(s), rtol=1e-15, atol=1e-17)
    row = [0.0, 3.0, 2.0, 0.1, 0.2, 0.0584, 0.7, 0.3, 0.5, 0.02]
    np.testing.assert_allclose(capi.initialize_state(row[6], row[3], row[4], row[5], row[7], row[8], row[9]), po.initialize_state(row)[:21],
                               rtol=1e-15, atol=1e-17)


def test_synth_batch_matches_initialize_state():
    x0, ctrl, gt = synth.make_batch(5, 7, seed=11)
    assert x0.shape == (21, 5) and ctrl.shape == (6, 2, 5) and gt.shape == (7, 3, 5)
    v0 = synth.B @ ctrl[0]
    for n in range(5):
        row = [0.0, ctrl[0, 0, n], ctrl[0, 1, n], 0.0, 0.0, synth.Z_STABLE_NN, gt[0, 0, n], v0[2, n], v0[0, n], v0[1, n]]
        np.testing.assert_allclose(x0[:, n], po.initialize_state(row)[:21], rtol=1e-14, atol=1e-16)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_gpu_fails_loudly():
    with pytest.raises(capi.AvsError):
        capi.Context(0, capi.TIRE_NN)


def test_legacy_net_csv_parser_handles_the_reference_file_quirks():
    """data/nn_pretrained.csv: 435 numbers on 434 lines, one line carrying the junk token ';in_mean' between two numbers
    (SURVEY.md 0.3).  The fixture reproduces those quirks with its own numbers; the real file is checked where it is mounted."""
    import os
    import numpy as np
    from auvsl_neural_ode_b200 import capi
    here = os.path.dirname(os.path.abspath(__file__))
    q = capi.load_legacy_net_csv(os.path.join(here, "golden", "legacy_net_like.csv"))
    assert np.array_equal(q, np.load(os.path.join(here, "golden", "legacy_net_like.npy")))
    ref = "/root/reference/data/nn_pretrained.csv"
    if os.path.exists(ref):
        r = capi.load_legacy_net_csv(ref)
        assert r.size == 435 and r[0] == 1.2790e+02 and r[424] == 285.0602943015066 and r[425] == 0.0 and r[434] == 0.028547002200876626
    import pytest
    with pytest.raises(capi.AvsError):
        capi.load_legacy_net_csv(os.path.join(here, "golden", "legacy_net_like.npy"))


def test_committed_counters_belong_to_the_committed_kernel_sources():
    """bench.py builds its roofline blocks from profiles/r2_counters.json; the capture must be of THIS csrc/ (a kernel change
    without a fresh ncu capture would otherwise go unnoticed until `counters.matches_build` shows false on the GPU box)"""
    import json
    import bench
    c = json.load(open(os.path.join(ROOT, "profiles", "r2_counters.json")))
    assert c["source_hash"] == bench.source_hash()
    assert {"rollout_fwd_kernel<NN,flat,STORE>", "rollout_bwd2_kernel<flat>", "rollout_fwd_kernel<NN,flat>", "rollout_fwd_kernel<Bekker,dyn>"} <= set(c["kernels"])
