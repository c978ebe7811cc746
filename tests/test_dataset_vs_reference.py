"""Window supplier (zeggs_b200.data.WindowDataset) against the LIVE reference dataset (ZEGGS/dataset.py:9-204) on a synthetic
processed_data.npz with the reference's schema (data_pipeline.py:650-684): same sliding windows, same style-example windows
(edge clamping, tail repetition), same one-hot labels.  Skipped where the reference tree is absent (the GPU box)."""
import json

import numpy as np
import pytest
import torch

from oracle import ref_shim
from zeggs_b200 import synth
from zeggs_b200.data import KEYS, WindowDataset

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("ds")
    st = synth.load_stats()
    rs = np.random.RandomState(3)
    N = 900
    data = {"X_audio_features": rs.randn(N, 81).astype(np.float32)}
    win = synth.make_pose_windows(1, N, seed=4)
    for k in KEYS:
        data["Y_" + k] = win[k][0]
    ranges = np.array([[0, 300], [300, 420], [420, 900]], dtype=np.int64)       # a short range exercises the clamping
    data.update(ranges_train=ranges, ranges_valid=ranges[:1], ranges_train_labels=np.array([0, 2, 1]),
                ranges_valid_labels=np.array([0]))
    for k in ("audio_input_mean", "audio_input_std", "anim_input_mean", "anim_input_std", "anim_output_mean", "anim_output_std"):
        data[k] = st[k]
    np.savez(d / "processed_data.npz", **data)
    details = dict(bone_names=[f"b{i}" for i in range(75)], label_names=["Neutral", "Happy", "Sad"],
                   parents=[int(p) for p in st["parents"]], dt=float(st["dt"]))
    with open(d / "data_definition.json", "w") as f:
        json.dump(details, f)
    return d / "data_definition.json", d / "processed_data.npz"


@pytest.mark.parametrize("window,ex_len", [(64, 128), (100, 256), (64, 64)])
def test_windows_and_examples_match_reference(files, window, ex_len):
    ref_shim.install()
    import dataset as ref_dataset
    ddef, dproc = files
    ref = ref_dataset.SGDataset(ddef, dproc, window, "example", ex_len)
    ours = WindowDataset(ddef, dproc, window, "example", ex_len, seed=0)
    assert len(ours) == len(ref)
    rs = np.random.RandomState(window)
    for i in list(rs.randint(0, len(ref), size=40)) + [0, len(ref) - 1]:
        item = ref[int(i)]
        start, ri = int(ours.starts[i]), int(ours.rng_idx[i])
        rows = torch.arange(start, start + window)
        assert torch.equal(rows, ref.R[int(i)]) and ri == int(ref.S[int(i)])
        assert torch.equal(ours.X[rows], item[0])
        for j, k in enumerate(KEYS):
            assert torch.equal(ours.Y[k][rows], item[1 + j]), k
        ex = ours._example(start, ri)
        assert ex.shape == item[10].shape and torch.equal(ex, item[10])


def test_label_batches_and_host_batches(files):
    ref_shim.install()
    import dataset as ref_dataset
    ddef, dproc = files
    ref = ref_dataset.SGDataset(ddef, dproc, 64, "label", 128)
    ours = WindowDataset(ddef, dproc, 64, "label", 128, seed=5)
    b = ours.sample_host_batch(6)
    assert b["style"].shape == (6, 3) and torch.all(b["style"].sum(1) == 1)
    assert b["audio"].shape == (6, 64, 81) and b["lpos"].shape == (6, 64, 75, 3)
    # every sampled label is the label of the range its window came from
    ours2 = WindowDataset(ddef, dproc, 64, "label", 128, seed=5)
    idx = ours2.rs.randint(0, len(ours2.starts), size=6)
    for r, i in enumerate(idx):
        assert torch.equal(b["style"][r], ref.L[int(i)])
