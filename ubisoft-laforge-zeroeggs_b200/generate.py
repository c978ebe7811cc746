"""Inference on the B200 path with the reference's call surface (ZEGGS/generate.py:22-37, 411).

`generate_motion()` is the accelerated core (arrays in, pose tensors out): mel front end -> SpeechEncoder ->
StyleEncoder -> persistent autoregressive decoder, all in libzeggs_b200.so.  `generate_gesture()` keeps the
reference signature; its file-format edges (BVH parse, `preprocess_animation`, BVH write -- SURVEY.md §2 rows 11/13,
out of the accelerated scope) are delegated to the reference's own helpers when they are importable (i.e. when this
runs inside the reference tree as a drop-in), and otherwise pre-processed arrays are accepted / an .npz is written.
"""
import json
import pathlib
import sys
from pathlib import Path
from shutil import copyfile

import numpy as np
import torch

from . import _lib, audio, modules, ops

POSE_KEYS = ["root_pos", "root_rot", "root_vel", "root_vrt", "lpos", "ltxy", "lvel", "lvrt"]


def load_networks(network_path, device, with_style=True):
    """torch.load of the whole-module pickles (generate.py:130-138); classes named `modules.*` resolve to ours."""
    saved = sys.modules.get("modules")
    sys.modules["modules"] = modules
    try:
        nets = {}
        for n in ("speech_encoder", "decoder") + (("style_encoder",) if with_style else ()):
            nets[n] = torch.load(Path(network_path) / f"{n}.pt", map_location="cpu", weights_only=False).to(device).eval()
    finally:
        if saved is not None:
            sys.modules["modules"] = saved
        else:
            del sys.modules["modules"]
    return nets


def read_wav(path):
    """scipy read + the reference's int->float rescale (audio_files.py:211-236); 16 kHz mono expected (generate.py:161-168)."""
    from scipy.io import wavfile
    fs, x = wavfile.read(str(path))
    if x.ndim > 1:
        x = x[:, 0]
    if x.dtype == np.int16:
        x = x / 32768.0
    elif x.dtype == np.int32:
        x = x / 2147483648.0
    elif x.dtype == np.uint8:
        x = ((x / 255.0) - 0.5) * 2
    if fs != 16000:
        raise _lib.ZeggsError(f"{path}: expected 16 kHz audio, got {fs} Hz (resample first; the reference shells out to SoX)")
    return x.astype(np.float32)


@torch.no_grad()
def generate_motion(nets, stats, audio_conf, audio_data, style, first_pose, gaze_pos0, dt, temperature=1.0, eps=None,
                    device="cuda"):
    """audio_data [n_samples] (or [N, n_samples]) float32 @16 kHz; style: normalised-able example [T_ex,1134] (raw) or an
    embedding [Z]; first_pose: dict of the 8 first-frame tensors.  Returns the decoder's 8-tuple (batch N)."""
    dev = torch.device(device)
    f = lambda k: torch.as_tensor(stats[k], dtype=torch.float32, device=dev)
    wav = torch.as_tensor(audio_data, dtype=torch.float32, device=dev)
    if wav.dim() == 1:
        wav = wav[None]
    N = wav.shape[0]
    n_frames = int(round(60.0 * (wav.shape[1] / 16000)))                                  # generate.py:170
    feats = audio.preprocess_audio(wav, 60, n_frames, audio_conf, ["mel_spec", "energy"], device=dev)
    if feats.dim() == 2:
        feats = feats[None]
    speech = nets["speech_encoder"]((feats - f("audio_input_mean")) / f("audio_input_std"))
    style = torch.as_tensor(style, dtype=torch.float32, device=dev)
    if style.dim() == 2 and style.shape[-1] == modules.P_IN:                              # a raw example: encode it
        ex = (style[None] - f("anim_input_mean")) / f("anim_input_std")
        z, _, _ = nets["style_encoder"](ex, temperature, eps=eps if eps is not None else None)
    else:
        z = style.reshape(1, -1)
    z = z.expand(N, -1)
    T = speech.shape[1]
    fp = {k: torch.as_tensor(first_pose[k], dtype=torch.float32, device=dev) for k in POSE_KEYS}
    fp = {k: (v[None] if v.dim() == {"root_pos": 1, "root_rot": 1, "root_vel": 1, "root_vrt": 1, "lpos": 2, "ltxy": 3, "lvel": 2, "lvrt": 2}[k] else v)
          for k, v in fp.items()}
    fp = {k: v.expand(N, *v.shape[1:]) for k, v in fp.items()}
    gaze = torch.as_tensor(gaze_pos0, dtype=torch.float32, device=dev).reshape(1, 1, 3).expand(N, T, 3)   # generate.py:374-376
    return nets["decoder"](*[fp[k] for k in POSE_KEYS], gaze, speech, z.unsqueeze(1).repeat(1, T, 1), None,
                           f("anim_input_mean"), f("anim_input_std"), f("anim_output_mean"), f("anim_output_std"), float(dt)), z


def _reference_io():
    """The reference's own file-format helpers, if this process can import them (drop-in use inside the reference tree)."""
    try:
        from anim import bvh, quat                      # noqa: F401
        from data_pipeline import preprocess_animation  # noqa: F401
        from utils import write_bvh                     # noqa: F401
        return dict(bvh=bvh, quat=quat, preprocess_animation=preprocess_animation, write_bvh=write_bvh)
    except Exception:
        return None


def generate_gesture(audio_file, styles, network_path, data_path, results_path, style_encoding_type="example",
                     blend_type="add", blend_ratio=[0.5, 0.5], file_name=None, first_pose=None, temperature=1.0,
                     seed=1234, use_gpu=True, use_script=False):
    """Drop-in for ZEGGS/generate.py:22.  Same arguments and return value (the final style encoding)."""
    assert (audio_file is None) == (results_path is None)                                # generate.py:84
    if not (use_gpu and torch.cuda.is_available()):
        raise _lib.ZeggsError("zeggs_b200.generate_gesture needs a CUDA device (no CPU fallback)")
    np.random.seed(seed); torch.manual_seed(seed)
    device = torch.device("cuda:0")
    data_path, network_path = Path(data_path), Path(network_path)
    conf = json.load(open(data_path / "data_pipeline_conf.json"))
    details = json.load(open(data_path / "data_definition.json"))
    stats = dict(np.load(data_path / "stats.npz"))
    nets = load_networks(network_path, device, with_style=(style_encoding_type == "example"))
    io = _reference_io()
    label_names, dt = details["label_names"], details["dt"]
    f = lambda k: torch.as_tensor(stats[k], dtype=torch.float32, device=device)

    def animation(src):
        if isinstance(src, dict):
            return src
        if io is None:
            raise _lib.ZeggsError("BVH parsing / preprocess_animation are outside the accelerated path: run inside the reference "
                                  "tree (its anim.bvh + data_pipeline are used) or pass pre-processed arrays")
        keys = ["root_pos", "root_rot", "root_vel", "root_vrt", "lpos", "lrot", "ltxy", "lvel", "lvrt", "cpos", "crot", "ctxy",
                "cvel", "cvrt", "gaze_pos", "gaze_dir"]
        return dict(zip(keys, io["preprocess_animation"](io["bvh"].load(src) if not isinstance(src, dict) else src)))

    encs, last_anim = [], None
    with torch.no_grad():
        for style in styles:
            if style_encoding_type == "label":
                e = torch.zeros((1, len(label_names)), device=device); e[0, label_names.index(style)] = 1.0
            elif isinstance(style[0], np.ndarray):
                e = torch.as_tensor(style[0], dtype=torch.float32, device=device)[None]
            else:
                a = animation(style[0]); last_anim = a
                sl = slice(*style[1]) if style[1] is not None else slice(None)
                n = len(a["root_vel"][sl])
                vec = np.concatenate([a[k][sl].reshape(n, -1) for k in ("root_vel", "root_vrt", "lpos", "ltxy", "lvel", "lvrt")]
                                     + [np.zeros((n, 3), np.float32)], axis=1)                # gaze slot zero (generate.py:240-251)
                ex = (torch.as_tensor(vec, dtype=torch.float32, device=device) - f("anim_input_mean")) / f("anim_input_std")
                e, _, _ = nets["style_encoder"](ex[None], temperature)
            encs.append(e)
        if len(encs) > 1 and blend_type == "add":
            final = torch.matmul(torch.stack(encs, dim=1).transpose(2, 1), torch.tensor(blend_ratio, device=device))
        else:
            final = encs[0]
        if audio_file is None:
            return final
        a = animation(first_pose) if first_pose is not None else last_anim
        if a is None:
            raise _lib.ZeggsError("first_pose is required when no style example provides one (generate.py:313-354)")
        fp = {k: np.asarray(a[k][0]) for k in POSE_KEYS}
        wav = read_wav(audio_file)
        ac = conf["audio_conf"] if "audio_conf" in conf else conf
        ac = dict(ac, normalize_loudness=False)    # the BS.1770 gain (pyloudnorm) is outside the accelerated path
        out, _ = generate_motion(nets, stats, ac, wav, final[0], fp, np.asarray(a["gaze_pos"][0]), dt, temperature, device=device)
    V = {k: v[0].cpu().numpy() for k, v in zip(POSE_KEYS, out)}
    results_path = Path(results_path); results_path.mkdir(parents=True, exist_ok=True)
    file_name = file_name or f"audio_{Path(audio_file).stem}"
    if io is not None:
        from anim.txform import xform_orthogonalize_from_xy
        lrot = io["quat"].from_xform(xform_orthogonalize_from_xy(torch.as_tensor(V["ltxy"])).numpy())
        io["write_bvh"](str(results_path / (file_name + ".bvh")), V["root_pos"], V["root_rot"], V["lpos"], lrot,
                        parents=np.asarray(details["parents"]), names=details["bone_names"], order="zyx", dt=dt,
                        start_position=np.array([0, 0, 0]), start_rotation=np.array([1, 0, 0, 0]))
    else:
        np.savez(results_path / (file_name + ".npz"), **V)
    copyfile(audio_file, str(results_path / (file_name + ".wav")))
    return final
