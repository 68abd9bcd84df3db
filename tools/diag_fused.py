"""Diagnostics for the fused block kernel (run on the GPU box): error statistics instead of asserts."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cfdbench_b200 import synth, _lib
from oracle import fno_numpy as onp, fno_torch_port as opt
import test_gpu_fused as tf
from test_gpu_parity import dev, load_case, make_model, rel, stream
lib = _lib.load()

for batch in (1, 3, 80):
    rng = np.random.default_rng(20 + batch)
    ym = ((rng.standard_normal((batch, 32, 24, 12)) + 1j * rng.standard_normal((batch, 32, 24, 12))) * 40.0).astype(np.complex64)
    x = torch.from_numpy(rng.standard_normal((batch, 32, 64, 64)).astype(np.float32)).to(torch.bfloat16)
    w0 = (rng.standard_normal((32, 32)) / 6).astype(np.float32)
    bias = rng.standard_normal(32).astype(np.float32)
    img = torch.from_numpy(tf.encode_ym_image(ym)).cuda()
    xd, w0td, biasd = x.cuda(), dev(w0.T.copy()), dev(bias)
    out = torch.zeros(batch, 32, 64, 64, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.fno_block_fused(img.data_ptr(), xd.data_ptr(), w0td.data_ptr(), biasd.data_ptr(), out.data_ptr(), batch, stream()), "fused")
    torch.cuda.synchronize()
    spec = onp.spectral_inverse(ym.astype(np.complex128), 64, 64, 12, 12)
    lin = spec + np.einsum("oi,bihw->bohw", w0.astype(np.float64), x.float().numpy().astype(np.float64)) + bias.astype(np.float64)[None, :, None, None]
    ref = onp.gelu(lin)
    got = out.float().cpu().numpy().astype(np.float64)
    ref16 = torch.from_numpy(ref.astype(np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)
    d = np.abs(got - ref16)
    ulp = tf.bf16_ulp(ref16)
    bad = d > 1.0001 * ulp + 2e-6 * np.maximum(1, np.abs(lin))
    print(f"B={batch}: rel {rel(got, ref):.3e} flips {(d>0).mean():.3e} max|d| {d.max():.3e} max ulps {(d/ulp).max():.1f} "
          f"bad(abs-aware) {bad.sum()} |lin|max {np.abs(lin).max():.2f}")
    if bad.sum():
        idx = np.argwhere(bad)
        print("   first bad:", idx[:8].tolist())
        bs, os_, hs, ws = idx.T
        print("   bad per sample", np.bincount(bs, minlength=batch)[:16], "per h", np.bincount(hs, minlength=64), "per o", np.bincount(os_, minlength=32))
        i0 = tuple(idx[0]); print("   got", got[i0], "ref16", ref16[i0], "lin", lin[i0])

for name in ("cavity_b2_gain200", "cylinder_b2_gain200"):
    g, sd, batch, p = load_case(name)
    tb = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
    pp = opt.params_from_numpy(sd)
    cb = {k: torch.from_numpy(v) for k, v in batch.items()}
    with torch.no_grad():
        o16 = opt.forward(pp, cb["inputs"], cb["case_params"], cb["mask"], round_fn=opt.bf16_round, return_acts=True)
    res = {}
    for fused in (True, False):
        m = make_model(sd, p, act_dtype="bfloat16"); m.fused_block = fused
        with torch.no_grad():
            res[fused] = m(**tb)["preds"].cpu().numpy()
        # per-layer activations through the C ABI workspace
    print(name, "fused vs oracle16 %.3e  unfused vs oracle16 %.3e  fused vs unfused %.3e  fused vs fp32 golden %.3e  unfused vs golden %.3e" % (
        rel(res[True], o16["preds"].numpy()), rel(res[False], o16["preds"].numpy()), rel(res[True], res[False]),
        rel(res[True], g["preds"]), rel(res[False], g["preds"])))
