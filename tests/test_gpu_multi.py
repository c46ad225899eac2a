"""GPU suite, multi-GPU part (skipped below 2 devices): the data-parallel training step with the all-reduce behind the C ABI
(avs_comm_* / avs_allreduce_dev), from both host languages.
  * C++ host: `synth_demo --gpus 2` (Trainer::trainThreads -> one context + one host thread per GPU, avs_comm_init_all)
  * Python host: `scripts/dp_train_csv.py` under torchrun (one process per GPU, avs_comm_init_rank)
Both must leave bit-identical parameters on every replica and agree with each other bit for bit; the single-GPU C++ run
(host RMSProp, different summation order) must agree to rounding."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "auvsl_neural_ode_b200", "host")


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _fingerprint(text):
    m = re.search(r"params sum ([-+0-9.eE]+) weighted ([-+0-9.eE]+) p\[0\] ([-+0-9.eE]+) p\[103\] ([-+0-9.eE]+)", text)
    assert m, text[-2000:]
    return [float(x) for x in m.groups()]


@pytest.mark.skipif(_device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_training_cpp_and_python_hosts_agree(tmp_path):
    d = str(tmp_path) + "/"
    demo = os.path.join(HOST, "examples", "synth_demo")
    one = subprocess.run([demo, d, "--gpus", "1"], capture_output=True, text=True, timeout=900)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    os.remove(d + "tire.net")  # every run starts from the default parameters
    two = subprocess.run([demo, d, "--gpus", "2"], capture_output=True, text=True, timeout=900)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-2000:]
    assert "GPUs: 2" in two.stdout
    f1, f2 = _fingerprint(one.stdout), _fingerprint(two.stdout)
    np.testing.assert_allclose(f2, f1, rtol=1e-11)
    avg1 = [float(x) for x in re.findall(r"Avg Loss: ([-+0-9.eE]+)", one.stdout)]
    avg2 = [float(x) for x in re.findall(r"Avg Loss: ([-+0-9.eE]+)", two.stdout)]
    assert len(avg2) == 3 and avg2 == pytest.approx(avg1, rel=1e-5)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    py = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                         "--master-port", "29641", os.path.join(ROOT, "scripts", "dp_train_csv.py"), d, "3"],
                        capture_output=True, text=True, timeout=900, env=env)
    assert py.returncode == 0, py.stdout[-3000:] + py.stderr[-3000:]
    assert "ranks_bit_identical=True" in py.stdout and "native_comm=True" in py.stdout
    fp = _fingerprint(py.stdout)
    print("C++ --gpus 1:", f1, "\nC++ --gpus 2:", f2, "\nPython 2 ranks:", fp)
    assert fp == f2, "C++ host (threads + avs_comm_init_all) and Python host (processes + avs_comm_init_rank) must agree bit for bit"
