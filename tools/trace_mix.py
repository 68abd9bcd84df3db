"""Timeline of mode_mix_tc_kernel's roles (CTA 0) from the -DFNO_FZ_TRACE variant of the library (tools/trace_fused.py
--build-only builds it): one launch at B=256 (image output), per-tile intervals.  Usage (GPU box): python tools/trace_mix.py"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "cfdbench_b200", "build", "trace", "libtrace.so")
if not os.path.exists(so):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "trace_fused.py"), "--build-only"])
import numpy as np, torch
from cfdbench_b200 import _lib
_lib.LIB_PATH = so
lib = _lib.load()
lib.fno_debug_mix_trace.argtypes = [C.c_void_p]
batch = int(os.environ.get("B", "256"))
NT = 4 * 16 * 8
trace = torch.zeros(NT + 148 * 8, dtype=torch.int64, device="cuda")
xm = torch.randn(288, batch, 32, 2, device="cuda")
wop = torch.randn(288 * 8192, device="cuda")
img = torch.empty(batch * 147456, dtype=torch.uint8, device="cuda")
ym = torch.empty(288, batch, 32, 2, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def launch(image=True):
    if image:
        _lib.check(lib.fno_mode_mix_image(xm.data_ptr(), wop.data_ptr(), img.data_ptr(), batch, st), "mix image")
    else:
        _lib.check(lib.fno_mode_mix(xm.data_ptr(), wop.data_ptr(), ym.data_ptr(), batch, st), "mix")
for it in range(3):
    if it == 2:
        assert lib.fno_debug_mix_trace(trace.data_ptr()) == 0
    launch()
    torch.cuda.synchronize()
assert lib.fno_debug_mix_trace(None) == 0
for image in (True, False):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, z in ev:
        a.record(); launch(image); z.record()
    torch.cuda.synchronize()
    print("standalone kernel (%s output), trace off: us per launch" % ("image" if image else "ym"), sorted(round(a.elapsed_time(z) * 1e3, 1) for a, z in ev))
raw = trace.cpu().numpy()
cta = raw[NT:NT + 148 * 4].reshape(148, 4)
gt = raw[NT + 148 * 4:].reshape(148, 4)
print("globaltimer: first CTA start -> last CTA end %d ns; CTA durations ns min %d median %d max %d; SM clock %.2f GHz" % (
    gt[:, 2].max() - gt[:, 0].min(), (gt[:, 2] - gt[:, 0]).min(), np.median(gt[:, 2] - gt[:, 0]), (gt[:, 2] - gt[:, 0]).max(),
    float(np.median((cta[:, 2] - cta[:, 0]) / np.maximum(gt[:, 2] - gt[:, 0], 1)))))
print("per-CTA prologue cycles median %d; total cycles min %d median %d max %d" % (np.median(cta[:, 1] - cta[:, 0]), (cta[:, 2] - cta[:, 0]).min(), np.median(cta[:, 2] - cta[:, 0]), (cta[:, 2] - cta[:, 0]).max()))
t = raw[:NT].reshape(4, 16, 8)
rel = np.where(t > 0, t - cta[0, 0], -1)
for role, nev, name in ((2, 1, "producer: TMA issued"), (3, 5, "MMA issue: start, A landed (+D free), 16 issued, lo ready, all issued"),
                        (0, 4, "converter warp 0: start, A landed, lo buffer free, lo written"), (1, 3, "epilogue warp 8: start, D full, stored")):
    print(name)
    for i in range(16):
        r = rel[role, i]
        if r[0] < 0: continue
        print(f"  t{i:2d} " + " ".join(f"{v:7d}" for v in r[:nev]) + "   d: " + " ".join(f"{r[k+1]-r[k]:6d}" for k in range(nev - 1)))
