"""Host-side data-format edges of generate_gesture (CPU): BVH parse / feature extraction / BVH write of this package against the
reference's own functions (when /root/reference is importable, i.e. in the dev container) and against the committed goldens."""
import copy
import os

import numpy as np
import pytest

from oracle import pose_oracle as po, ref_shim
from tests import _fixtures as fx
from zeggs_b200 import animation, bvhio, generate

needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def test_pose_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pose_post.npz"))
    for n in range(2):
        pos, eul = po.pose_to_bvh_channels(g["root_pos"][n], g["root_rot"][n], g["lpos"][n], g["ltxy"][n])
        assert np.abs(pos - g[f"positions{n}"]).max() <= 2e-5 * max(1.0, np.abs(g[f"positions{n}"]).max())
        assert np.abs(eul - g[f"rotations{n}"]).max() <= 2e-3          # degrees
        q = po.quat_from_xform(po.orthogonalize_from_xy(g["ltxy"][n]))
        assert np.abs(q - g[f"lrot{n}"]).max() <= 1e-5


def test_split_by_ratio():
    assert generate.split_by_ratio(240, [0.5, 0.5]) == [[0, 120], [120, 240]]
    assert generate.split_by_ratio(241, [0.25, 0.75]) == [[0, 60], [60, 241]]


def test_bvh_round_trip(tmp_path):
    p = fx.make_synthetic_bvh(str(tmp_path / "s.bvh"), frames=12)
    a = animation.load_bvh(p)
    assert a["rotations"].shape == (12, 75, 3) and a["order"] == "zyx" and abs(a["frametime"] - 0.016667) < 1e-9
    bvhio.save_bvh(str(tmp_path / "t.bvh"), a["positions"], a["rotations"], a["parents"], a["names"], a["order"], a["frametime"])
    b = animation.load_bvh(str(tmp_path / "t.bvh"))
    assert np.abs(a["rotations"] - b["rotations"]).max() <= 1e-5 and np.abs(a["positions"] - b["positions"]).max() <= 1e-5
    assert list(a["parents"]) == list(b["parents"]) and a["names"] == b["names"]


@needs_ref
def test_bvh_writer_is_byte_identical_to_the_reference_writer(tmp_path):
    ref_shim.install()
    from anim import bvh
    d = fx.skeleton()
    rs = np.random.RandomState(0)
    pos = rs.randn(5, 75, 3).astype(np.float32); rot = (rs.randn(5, 75, 3) * 50).astype(np.float32)
    bvh.save(str(tmp_path / "a.bvh"), dict(order="zyx", offsets=pos[0], names=d["bone_names"], frametime=d["dt"],
                                           parents=np.array(d["parents"]), positions=pos, rotations=rot))
    bvhio.save_bvh(str(tmp_path / "b.bvh"), pos, rot, d["parents"], d["bone_names"], "zyx", d["dt"])
    assert open(tmp_path / "a.bvh").read() == open(tmp_path / "b.bvh").read()


@needs_ref
def test_load_and_preprocess_animation_match_the_reference(tmp_path):
    ref_shim.install()
    from anim import bvh
    from data_pipeline import preprocess_animation as ref_pa
    p = fx.make_synthetic_bvh(str(tmp_path / "s.bvh"), frames=120)
    a_ref, a = bvh.load(p), animation.load_bvh(p)
    for k in ("rotations", "positions", "offsets", "parents"):
        assert np.abs(np.asarray(a_ref[k], dtype=np.float64) - np.asarray(a[k], dtype=np.float64)).max() == 0.0
    keys = ["root_pos", "root_rot", "root_vel", "root_vrt", "lpos", "lrot", "ltxy", "lvel", "lvrt", "cpos", "crot", "ctxy", "cvel",
            "cvrt", "gaze_pos", "gaze_dir"]
    for cut in (None, (10, 80)):
        ar = copy.deepcopy(a_ref)
        if cut:
            ar["rotations"], ar["positions"] = ar["rotations"][cut[0]:cut[1]], ar["positions"][cut[0]:cut[1]]
        R = dict(zip(keys, ref_pa(ar)))
        O = animation.preprocess_animation(animation.trim(a, cut))
        for k, v in O.items():
            assert np.abs(R[k].astype(np.float64) - v).max() <= 1e-5 * max(1.0, np.abs(R[k]).max()), k
