"""Generates tests/golden/batch.npz by RUNNING THE REAL REFERENCE batch assembly (utils_motion_vae.py: rand_rotation_matrix,
change_fps, MotionSeqData.__getitem__) on a small synthetic AMASS-layout file.  Authoring container only:

    python oracle/make_golden_batch.py

The reference draws its crop offset / fps / rotation from the global `random` and `np.random` generators; the script seeds them,
runs `__getitem__`, then re-seeds and replays the same draws to record WHICH window and WHICH three random numbers were used, so
that the oracle (oracle/batch_ref.py) and the CUDA kernel can be checked on exactly the same inputs.
"""
import json
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from make_golden import REF, import_reference  # noqa: E402


def main():
    import_reference()
    import utils_motion_vae as U  # the reference module, unmodified

    rng = np.random.RandomState(7)
    mean_std = np.load(f"{REF}/utils/data/for_all_data_motion_model/all_amass_data_mean_std.npy")
    out = {"mean_std_dtype": np.array(str(mean_std.dtype))}
    # ---- rand_rotation_matrix known answers
    rn = rng.uniform(size=(6, 3))
    out["rr_randnums"] = rn
    out["rr_full"] = np.stack([U.rand_rotation_matrix(1.0, r) for r in rn])
    out["rr_small"] = np.stack([U.rand_rotation_matrix(0.25, r) for r in rn])
    # ---- a dataset of 2 synthetic sequences in the 579-column layout (valid rotations in the 6D / matrix columns)
    tmp = tempfile.mkdtemp()
    seqs = []
    for i, T_total in enumerate((40, 96)):
        x6 = torch.from_numpy(rng.randn(T_total, 24, 6)).float()
        import my_tools
        R = my_tools.rotation_matrix_from_ortho6d(x6.view(-1, 6)).view(T_total, 24, 3, 3)
        six = torch.stack((R[..., 0], R[..., 1]), dim=-2).reshape(T_total, 144)
        rest = torch.from_numpy(rng.randn(T_total, 579 - 144 - 216)).float()
        seq = torch.cat((six, R.reshape(T_total, 216), rest), dim=1).numpy().astype(np.float32)
        np.save(os.path.join(tmp, "seq%d.npy" % i), seq)
        seqs.append(seq)
        out["seq%d" % i] = seq
    json.dump({"0": "seq0.npy", "1": "seq1.npy"}, open(os.path.join(tmp, "ids.json"), "w"))
    ms_path = os.path.join(tmp, "ms.npy")
    np.save(ms_path, mean_std)
    names = ["rot6d", "rotmat", "rot_pos", "joint_pos", "linear_v", "angular_v", "root_v"]
    cases = [("plain", 0, 8, False, False, 11), ("rot", 0, 8, False, True, 12), ("rot64", 1, 64, False, True, 13),
             ("fps_rot", 1, 8, True, True, 14)]
    for tag, idx, T, fps, rot, seed in cases:
        ds = U.MotionSeqData(tmp, os.path.join(tmp, "ids.json"), ms_path, {"train_seq_len": T}, fps_aug_flag=fps,
                             random_root_rot_flag=rot)
        random.seed(seed)
        np.random.seed(seed)
        res = ds[idx]
        # replay the draws (same order as __getitem__: change_fps, crop offset, rand_rotation_matrix)
        random.seed(seed)
        np.random.seed(seed)
        data = seqs[idx]
        freq = 1
        if fps:
            tries = 0
            while tries < 10:
                f = random.sample([1, 2, 3, 4, 5, 6, 8, 10, 12], 1)[0]
                tries += 1
                if data[0::f].shape[0] >= T:
                    freq = f
                    break
        sub = data[0::freq]
        t0 = random.sample(list(range(sub.shape[0] - T + 1)), 1)[0]
        rnd = np.random.uniform(size=(3,)) if rot else np.zeros(3)
        out[f"{tag}_meta"] = np.array([idx, T, freq, t0, int(rot)])
        out[f"{tag}_randnums"] = rnd
        for n, v in zip(names, res):
            out[f"{tag}_{n}"] = v.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "batch.npz"), **out)
    print("wrote batch.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
