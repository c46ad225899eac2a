"""GPU suite, robustness part: compute-sanitizer is closed on this pool, so the hand-rolled named-barrier hand-shake of the
reverse kernel and the record-stream indexing are exercised by (bounded) soak runs instead — many shapes around the tile /
CTA boundaries, every case twice and compared bit for bit (a lost hand-shake shows up as a hang -> pytest timeout, a stale
ring slot or an out-of-range record as a differing or non-finite result) — plus the one small case that touches every kernel
family (scripts/sanitize_case.py, the compute-sanitizer target)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from auvsl_neural_ode_b200 import capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_soak_reverse_kernel_handshake_shapes():
    ctx = capi.Context(0, capi.TIRE_NN)
    cases = 0
    for terr in (capi.FLAT, capi.BUMPY):
        ctx.set_terrain(terr)
        for N in (1, 3, 7, 8, 9, 31, 33, 64, 255, 257, 1000):
            for T, sub in ((2, 1), (2, 10), (3, 3), (17, 7), (40, 10)):
                x0, ctrl, gt = synth.make_batch(N, T, seed=1000 + cases)
                a = ctx.rollout_grad(x0, ctrl, gt, T, substeps=sub)
                b = ctx.rollout_grad(x0, ctrl, gt, T, substeps=sub)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.all(np.isfinite(a[2])), (terr, N, T, sub)
                cases += 1
    ctx.set_terrain(capi.FLAT)
    x0, ctrl, gt = synth.make_batch(4096, 200)
    ref = None
    for it in range(8):
        loss, red, _ = ctx.rollout_grad(x0, ctrl, gt, 200, per_sample=False)
        if ref is None:
            ref = (loss.copy(), red.copy())
        assert np.array_equal(loss, ref[0]) and np.array_equal(red, ref[1]), it
    ctx.close()
    print(f"soak: {cases} small cases x 2 + 8 full-size steps, bit-identical repeats")


@pytest.mark.timeout(300)
def test_sanitize_case_runs_clean():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sanitize_case.py")], capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sanitize case ok" in out.stdout and "True" in out.stdout


@pytest.mark.timeout(600)
def test_index_checked_build_runs_clean():
    """libavsim_debug.so (make -C auvsl_neural_ode_b200/csrc debug; -DAVS_DEBUG=1) checks every address formed on the state- and
    activation-record streams against the extent of its buffer and traps on a violation.  The all-kernels case and a set of
    awkward shapes (tail lanes, chunked batches) must run clean on it and give the production build's results bit for bit."""
    dbg = os.path.join(ROOT, "auvsl_neural_ode_b200", "libavsim_debug.so")
    if not os.path.exists(dbg):
        pytest.skip("libavsim_debug.so not built")
    env = dict(os.environ, AVS_LIB=dbg)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sanitize_case.py")], capture_output=True, text=True, timeout=280, env=env)
    assert out.returncode == 0 and "AVS_DEBUG" not in out.stdout + out.stderr, out.stdout[-2000:] + out.stderr[-2000:]
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from auvsl_neural_ode_b200 import capi, synth\n"
        "ctx = capi.Context(0, capi.TIRE_NN)\n"
        "acc = []\n"
        "for N, T, sub in ((1, 2, 1), (7, 3, 10), (33, 17, 7), (257, 9, 10), (600, 6, 10)):\n"
        "    x0, ctrl, gt = synth.make_batch(N, T, seed=N)\n"
        "    if N == 600: ctx.set_checkpoint_budget((((T - 1) * sub * 4 + 1) * 14 * 8 + (T - 1) * sub * 4 * 20 * 4 * 8) * 256 + 64)\n"
        "    loss, red, gps = ctx.rollout_grad(x0, ctrl, gt, T, substeps=sub)\n"
        "    acc.append(float(loss.sum()) + float(np.abs(gps).sum()))\n"
        "print('RESULT', repr(acc))\n")
    res = {}
    for tag, e in (("debug", env), ("production", dict(os.environ))):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=280, env=e)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[tag] = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0]
    assert res["debug"] == res["production"]


def test_two_contexts_on_one_device_do_not_race():
    """Every translation unit keeps ONE __constant__ model per device; two contexts with different weights used through the
    asynchronous *_dev entry points on their own streams must each see their own model (launches are serialised against the
    previous launch of another stream on the device, avs_capi.cu)."""
    import torch
    dev = torch.device("cuda:0")
    a, b = capi.Context(0, capi.TIRE_NN), capi.Context(0, capi.TIRE_NN)
    pa = capi.nn_default_params()
    pb = pa.copy()
    pb[:104] *= 1.05
    a.set_nn_params(pa)
    b.set_nn_params(pb)
    N, T = 2048, 30
    x0, ctrl, _ = synth.make_batch(N, T, seed=21)
    _, want_a = a.rollout(x0, ctrl, T, want_list=False)
    _, want_b = b.rollout(x0, ctrl, T, want_list=False)
    assert not np.array_equal(want_a, want_b)
    d_x0, d_ctrl = torch.from_numpy(x0).to(dev), torch.from_numpy(ctrl).to(dev)
    fa = [torch.zeros(21, N, dtype=torch.float64, device=dev) for _ in range(6)]
    fb = [torch.zeros(21, N, dtype=torch.float64, device=dev) for _ in range(6)]
    torch.cuda.synchronize()
    for i in range(6):  # interleaved, no host synchronisation in between
        a.rollout_dev(N, T, 10, d_x0.data_ptr(), d_ctrl.data_ptr(), 0, fa[i].data_ptr())
        b.rollout_dev(N, T, 10, d_x0.data_ptr(), d_ctrl.data_ptr(), 0, fb[i].data_ptr())
    a.sync(); b.sync()
    for i in range(6):
        assert np.array_equal(fa[i].cpu().numpy(), want_a), i
        assert np.array_equal(fb[i].cpu().numpy(), want_b), i
    a.close(); b.close()
