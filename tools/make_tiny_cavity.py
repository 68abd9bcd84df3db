"""A tiny on-disk cavity dataset in the reference's format (reference src/dataset/cavity.py:15-34,384-446):
<out>/cavity/{prop,bc,geo}/case<N>/{u.npy, v.npy, case.json}; u, v: (T, 64, 64) float32.
Smooth synthetic flow fields that keep changing frame to frame (so the stable-state cut-off never triggers)."""
import json, os, sys
import numpy as np


def make(out_dir: str, n_cases=(4, 3, 3), frames: int = 6, seed: int = 0) -> str:
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, 64), np.linspace(0, 1, 64), indexing="ij")
    idx = 0
    for sub, n in zip(("prop", "bc", "geo"), n_cases):
        for c in range(n):
            d = os.path.join(out_dir, "cavity", sub, f"case{c:04d}")
            os.makedirs(d, exist_ok=True)
            a, b, ph = rng.uniform(0.5, 2.0, 3)
            u = np.stack([np.sin(2 * np.pi * (a * xx + 0.07 * t)) * np.cos(np.pi * b * yy + ph) * (1 + 0.1 * t)
                          for t in range(frames)]).astype(np.float32)
            v = np.stack([-np.cos(2 * np.pi * (a * xx + 0.07 * t)) * np.sin(np.pi * b * yy + ph) * (1 + 0.1 * t)
                          for t in range(frames)]).astype(np.float32)
            np.save(os.path.join(d, "u.npy"), u)
            np.save(os.path.join(d, "v.npy"), v)
            with open(os.path.join(d, "case.json"), "w") as f:   # key order = collate_fn's tensor order (train_auto.py:44-50)
                json.dump({"vel_top": float(rng.uniform(1, 50)), "density": float(rng.uniform(1, 10)),
                           "viscosity": float(rng.uniform(1e-3, 1e-2)), "height": 1.0, "width": 1.0}, f)
            idx += 1
    return out_dir


if __name__ == "__main__":
    print(make(sys.argv[1]))
