"""Golden windows from the reference's own dataset class (run in the build container only).

    python tests/golden/make_dataset_golden.py        # needs /root/reference

Imports WindowedAllMembersDataset_random (dataset_single_member.py:5-196, the class train.py:990
instantiates) and records what it returns for a small seeded (T, M, 1, H, W) array pair in its default
"consecutive" mode, for every decision the B200 data path reproduces:

  * no augmentation, no crop                      -> indexing idx -> (member, first frame, centre target)
  * time_reverse_p = 1 (always reversed)          -> the centre-preserving reversal (:178-185)
  * crop_mode = "center"                          -> crop origin (:157-159)
  * crop_mode = "random", np.random.seed(s)       -> the windows AND the crop origins the class drew,
                                                     recovered by locating the crop in the full frame

The arrays themselves are regenerated from the seed by the test, so the fixture holds outputs only.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from dataset_single_member import WindowedAllMembersDataset_random  # noqa: E402

T, M, H, W = 9, 3, 12, 20
SEED = 77


def arrays():
    rng = np.random.default_rng(SEED)
    cond = rng.standard_normal((T, M, 1, H, W)).astype(np.float32)
    tgt = rng.standard_normal((T, M, 1, H, W)).astype(np.float32)
    return cond, tgt


def collect(ds):
    cs, xs = zip(*(ds[i] for i in range(len(ds))))
    return np.stack([c.numpy() for c in cs]), np.stack([x.numpy() for x in xs])


def main():
    cond, tgt = arrays()
    out = {"shape": np.array([T, M, H, W]), "seed": np.array(SEED)}
    for K in (3, 4, 5):
        ds = WindowedAllMembersDataset_random(cond, tgt, K=K, center=True, crop_hw=None, time_reverse_p=0.0)
        out[f"plain_K{K}_cond"], out[f"plain_K{K}_x0"] = collect(ds)
        ds = WindowedAllMembersDataset_random(cond, tgt, K=K, center=True, crop_hw=None, time_reverse_p=1.0)
        out[f"rev_K{K}_cond"], out[f"rev_K{K}_x0"] = collect(ds)
    ds = WindowedAllMembersDataset_random(cond, tgt, K=3, center=True, crop_hw=(8, 10), crop_mode="center",
                                          time_reverse_p=0.0)
    out["center_crop_cond"], out["center_crop_x0"] = collect(ds)
    # oversize crop is clamped to the grid (:55-56)
    ds = WindowedAllMembersDataset_random(cond, tgt, K=3, center=True, crop_hw=(64, 10), crop_mode="center",
                                          time_reverse_p=0.0)
    out["clamped_crop_cond"], out["clamped_crop_x0"] = collect(ds)
    # random crop: values are unique floats, so the origin the class drew can be read back from x0
    np.random.seed(5)
    ds = WindowedAllMembersDataset_random(cond, tgt, K=3, center=True, crop_hw=(8, 10), crop_mode="random",
                                          time_reverse_p=0.0)
    c, x = collect(ds)
    origins = []
    for idx in range(len(ds)):
        m, t0 = idx % M, idx // M
        full = tgt[t0 + 1, m, 0]
        hit = np.argwhere(full == x[idx, 0, 0, 0])
        assert len(hit) == 1
        origins.append(hit[0])
    out["random_crop_cond"], out["random_crop_x0"], out["random_crop_origin"] = c, x, np.array(origins)
    path = os.path.join(os.path.dirname(__file__), "dataset_windows.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(ds), "items per case")


if __name__ == "__main__":
    main()
