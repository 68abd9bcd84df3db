"""Timeline of dft_fwd_tc_kernel's roles (CTA 0): uses the -DFNO_FZ_TRACE variant of the library that tools/trace_fused.py
builds into build/trace/, runs one launch at B=256 and prints per-batch intervals.  Usage (GPU box): python tools/trace_dft.py"""
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
lib.fno_debug_dft_trace.argtypes = [C.c_void_p]
batch = int(os.environ.get("B", "256"))
NT = 5 * 64 * 8
trace = torch.zeros(NT + 148 * 8, dtype=torch.int64, device="cuda")
x = torch.randn(batch, 32, 64, 64, device="cuda").bfloat16()
xm = torch.empty(288, batch, 32, dtype=torch.complex64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    if it == 2:
        assert lib.fno_debug_dft_trace(trace.data_ptr()) == 0
    _lib.check(lib.fno_spectral_dft_fwd_tc(x.data_ptr(), xm.data_ptr(), batch, 1.0, 1.0, st), "dft tc")
    torch.cuda.synchronize()
assert lib.fno_debug_dft_trace(None) == 0
def timed(label):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, z in ev:
        a.record()
        _lib.check(lib.fno_spectral_dft_fwd_tc(x.data_ptr(), xm.data_ptr(), batch, 1.0, 1.0, st), "dft tc")
        z.record()
    torch.cuda.synchronize()
    print(label, "median us", np.median([a.elapsed_time(z) * 1e3 for a, z in ev]).round(1))
lib.fno_debug_dft_knock.argtypes = [C.c_int]
for bits, label in ((0, "full kernel"), (1, "no stage-B MMAs"), (2, "no converter stores"), (4, "no mode stores"), (8, "no stage-A MMAs"), (3, "no B MMAs, no conv stores"),
                    (9, "no MMAs at all"), (11, "no MMAs, no conv stores"), (15, "waits + tmem reads only")):
    assert lib.fno_debug_dft_knock(bits) == 0
    timed("knock-out %2d %-28s" % (bits, label))
assert lib.fno_debug_dft_knock(0) == 0
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for a, z in ev:
    a.record()
    _lib.check(lib.fno_spectral_dft_fwd_tc(x.data_ptr(), xm.data_ptr(), batch, 1.0, 1.0, st), "dft tc")
    z.record()
torch.cuda.synchronize()
print("standalone kernel, trace off: us per launch", sorted(round(a.elapsed_time(z) * 1e3, 1) for a, z in ev))
n_cta = min(148, batch * 8)
cta = trace.cpu().numpy()[NT:NT + 148 * 4].reshape(148, 4)[:n_cta]
gt = trace.cpu().numpy()[NT + 148 * 4:].reshape(148, 4)[:n_cta]
print("globaltimer: first CTA start -> last CTA end %d ns; CTA starts spread %d ns; CTA durations ns min %d median %d max %d; SM clock %.2f GHz" % (
    gt[:, 2].max() - gt[:, 0].min(), gt[:, 0].max() - gt[:, 0].min(), (gt[:, 2] - gt[:, 0]).min(), np.median(gt[:, 2] - gt[:, 0]), (gt[:, 2] - gt[:, 0]).max(),
    float(np.median((cta[:, 2] - cta[:, 0]) / np.maximum(gt[:, 2] - gt[:, 0], 1)))))
pro, tot = cta[:, 1] - cta[:, 0], cta[:, 2] - cta[:, 0]
print("per-CTA prologue cycles: min %d median %d max %d;  total cycles: min %d median %d max %d" % (pro.min(), np.median(pro), pro.max(), tot.min(), np.median(tot), tot.max()))
t = trace.cpu().numpy()[:NT].reshape(5, 64, 8)
t0 = cta[0, 0]
rel = np.where(t > 0, t - t0, -1)
names = ["producer: start, slot free", "stage A issue: start, x landed, D_A free, issued", "converter warp 0: start, D_A full, drained, B2 free, staged",
         "stage B issue: start, B2 ready, D_B free, issued", "epilogue warp 16: start, D_B full, 8 q + drained, done"]
for role, nev in ((4, 2), (2, 4), (0, 5), (3, 4), (1, 4)):
    print(names[[4, 2, 0, 3, 1].index(role)])
    for i in range(20):
        r = rel[role, i]
        if r[0] < 0: continue
        print(f"  i{i:3d} " + " ".join(f"{v:7d}" for v in r[:nev]) + "   d: " + " ".join(f"{r[k+1]-r[k]:6d}" for k in range(nev - 1)))
