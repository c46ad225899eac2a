"""A/B timing of library variants built side by side (make OBJDIR=_obj_x TARGET=../libavsim_x.so EXTRA=-D...): each variant runs in
its own process through AVS_LIB.  usage: ab_variants.py tag1 tag2 ...   (tag 'main' = libavsim.so)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for tag in sys.argv[1:]:
    lib = os.path.join(ROOT, "auvsl_neural_ode_b200", "libavsim.so" if tag == "main" else f"libavsim_{tag}.so")
    env = dict(os.environ, AVS_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "quick_gpu.py"), "4096", "600", "200", "3"], capture_output=True, text=True, env=env).stdout
    lines = [l for l in out.splitlines() if l.startswith("grad=")]
    f600 = min(float(l.split("fwd_kernel=")[1].split()[0]) for l in lines if "grad=False" in l)
    fs = min(float(l.split("fwd_kernel=")[1].split()[0]) for l in lines if "grad=True" in l)
    bw = min(float(l.split("bwd_kernel=")[1].split()[0]) for l in lines if "grad=True" in l)
    print(f"{tag:10s} fwd-only(T=600) {f600:7.2f} ms   fwd+store(T=200) {fs:6.2f} ms   bwd {bw:6.2f} ms", flush=True)
