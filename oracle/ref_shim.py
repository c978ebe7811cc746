"""Import the UNMODIFIED reference (ZeroEGGS) from /root/reference for fixture generation.

TEST INFRASTRUCTURE ONLY.  Used solely by oracle/make_golden.py and by the `not gpu`
tests that pin the oracle restatement against the reference when /root/reference is
present (this dev container).  Nothing under `-m gpu`, `smoke()` or `bench.py` may
import this file: /root/reference does not exist on the GPU box.

The shims below touch no arithmetic:
  * `audio/__init__.py:11-21` raises on Linux without sox/ffmpeg  -> bare package stub
  * `scipy.signal.hann` (spectrograms.py:230) was removed         -> alias to windows.hann
  * `omegaconf`, `sox`, `pyloudnorm` not installed                -> empty stub modules
"""
import os
import sys
import types

REF_ROOT = os.environ.get("ZEGGS_REFERENCE_ROOT", "/root/reference")
REF_ZEGGS = os.path.join(REF_ROOT, "ZEGGS")


def available() -> bool:
    return os.path.isdir(REF_ZEGGS)


_installed = False


def install():
    """Put the reference on sys.path with the import shims; idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ZEGGS}")
    import scipy.signal
    import scipy.signal.windows

    if not hasattr(scipy.signal, "hann"):
        scipy.signal.hann = scipy.signal.windows.hann
    if "audio" not in sys.modules:
        pkg = types.ModuleType("audio")
        pkg.__path__ = [os.path.join(REF_ZEGGS, "audio")]
        sys.modules["audio"] = pkg
    for name in ("sox", "pyloudnorm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    try:
        import omegaconf  # noqa: F401
    except Exception:
        oc = types.ModuleType("omegaconf")

        class DictConfig(dict):
            def __getattr__(self, k):
                v = self[k]
                return DictConfig(v) if isinstance(v, dict) else v

        oc.DictConfig = DictConfig
        sys.modules["omegaconf"] = oc
    if REF_ZEGGS not in sys.path:
        sys.path.insert(0, REF_ZEGGS)
    _installed = True


def ref_modules():
    install()
    import modules  # the reference's ZEGGS/modules.py

    return modules


def ref_preprocess_audio():
    install()
    from data_pipeline import preprocess_audio  # ZEGGS/data_pipeline.py:33

    return preprocess_audio


def load_pretrained(version="v1"):
    """torch.load the shipped whole-module pickles (generate.py:130-138)."""
    import torch

    m = ref_modules()  # noqa: F841  (pickles resolve classes as `modules.*`)
    d = os.path.join(REF_ROOT, "data", "outputs", version, "saved_models")
    out = {}
    for n in ("speech_encoder", "decoder", "style_encoder"):
        p = os.path.join(d, n + ".pt")
        if os.path.exists(p):
            out[n] = torch.load(p, map_location="cpu", weights_only=False).eval()
    return out
