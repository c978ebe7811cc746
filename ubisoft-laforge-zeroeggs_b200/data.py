"""Window supplier with the reference's processed_data.npz schema (ZEGGS/data_pipeline.py:650-684,
ZEGGS/dataset.py:9-204): sliding training windows + the style-example window around each of them.
Whole arrays are kept pinned on the host and every batch is gathered with fancy indexing and copied with one
non-blocking H2D per tensor (the reference does 11 synchronous copies per step, train.py:215-225)."""
import json

import numpy as np
import torch

KEYS = ["root_pos", "root_rot", "root_vel", "root_vrt", "lpos", "ltxy", "lvel", "lvrt", "gaze_pos"]


class WindowDataset:
    def __init__(self, path_data_definition, path_processed_data, window, style_encoding_type, example_window_length, seed=0):
        with open(path_data_definition, "r") as f:
            self.details = json.load(f)
        d = np.load(path_processed_data)
        self.window = window
        self.style_encoding_type = style_encoding_type
        self.example_window_length = example_window_length
        self.nlabels = len(self.details["label_names"])
        self.ranges = d["ranges_train"]
        self.labels = d["ranges_train_labels"]
        self.X = torch.as_tensor(d["X_audio_features"], dtype=torch.float32)
        self.Y = {k: torch.as_tensor(d["Y_" + k], dtype=torch.float32) for k in KEYS}
        self.stats = {k: d[k] for k in ("audio_input_mean", "audio_input_std", "anim_input_mean", "anim_input_std",
                                        "anim_output_mean", "anim_output_std")}
        starts, rng_idx = [], []
        for i, (s, e) in enumerate(self.ranges):                       # dataset.py:82-93
            n = max(0, int(e) - window - int(s))
            starts.append(np.arange(int(s), int(s) + n))
            rng_idx.append(np.full(n, i))
        self.starts = np.concatenate(starts) if starts else np.zeros(0, np.int64)
        self.rng_idx = np.concatenate(rng_idx) if rng_idx else np.zeros(0, np.int64)
        self.rs = np.random.RandomState(seed)

    def __len__(self):
        return len(self.starts)

    def get_shapes(self):
        return dict(num_audio_features=self.X.shape[1], pose_input_size=len(self.stats["anim_input_std"]),
                    pose_output_size=len(self.stats["anim_output_std"]))

    def _example(self, start, ri):
        """dataset.py:176-204."""
        L, W = self.example_window_length, self.window
        s0, e0 = int(self.ranges[ri][0]), int(self.ranges[ri][1])
        first, last = start, start + W - 1
        ext = (L - W) // 2
        ws, we = min(ext, first - s0), min(ext, e0 - last)
        s_ext, w_ext = ws + ext - we, we + ext - ws
        a = max(first - s_ext, s0)
        b = min(min(last + w_ext, e0) + 1, len(self.Y["root_vel"]))
        n = b - a
        parts = [self.Y[k][a:b].reshape(n, -1) for k in ("root_vel", "root_vrt", "lpos", "ltxy", "lvel", "lvrt")]
        vec = torch.cat(parts + [torch.zeros(n, 3)], dim=1)
        if n < L:
            vec = torch.cat([vec, vec[-L + n:]], dim=0)
        return vec

    def sample_batch(self, batchsize, device):
        return {k: v.to(device, non_blocking=True) for k, v in self.sample_host_batch(batchsize).items()}

    def sample_host_batch(self, batchsize):
        """Same draw as sample_batch, left in pinned host memory (for DevicePrefetcher.upload)."""
        idx = self.rs.randint(0, len(self.starts), size=batchsize)
        rows = torch.as_tensor(self.starts[idx][:, None] + np.arange(self.window)[None, :])
        out = {"audio": self.X[rows]}
        for k in KEYS:
            out[k] = self.Y[k][rows]
        if self.style_encoding_type == "label":
            lab = torch.zeros(batchsize, self.nlabels)
            lab[torch.arange(batchsize), torch.as_tensor(self.labels[self.rng_idx[idx]]).long()] = 1.0
            out["style"] = lab
        else:
            out["style"] = torch.stack([self._example(int(self.starts[i]), int(self.rng_idx[i])) for i in idx])
        pin = torch.cuda.is_available()
        return {k: (v.pin_memory() if pin else v.contiguous()) for k, v in out.items()}


class DeviceWindowDataset(WindowDataset):
    """The same windows, drawn by the same generator, with the processed arrays RESIDENT IN HBM and every batch produced by ONE
    gather launch (zeggs_window_gather) -- no per-sample host indexing, no host->device copies beyond three int32 vectors of length
    B (window starts, example starts, example lengths).  SURVEY.md 8f row 2; replaces dataset.py:110-153, 176-204 and
    train.py:215-225.  sample_batch() returns the dict TrainStep.step() takes; bit-identical to WindowDataset.sample_host_batch()
    for the same seed (pure row copies)."""
    _EX = ("root_vel", "root_vrt", "lpos", "ltxy", "lvel", "lvrt")

    def __init__(self, *args, device="cuda", **kw):
        super().__init__(*args, **kw)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            from . import _lib
            raise _lib.ZeggsError("DeviceWindowDataset keeps the data in HBM: needs a CUDA device")
        self.names = ["audio"] + KEYS
        flat = [self.X] + [self.Y[k] for k in KEYS]
        self.dev_arrays = [a.reshape(a.shape[0], -1).contiguous().to(self.device) for a in flat]
        self.shapes = [tuple(a.shape[1:]) for a in flat]
        self.widths = [a.shape[1] for a in self.dev_arrays]
        self.n_frames = flat[0].shape[0]
        self.ex_width = sum(self.widths[self.names.index(k)] for k in self._EX) + 3

    def _example_range(self, start, ri):
        """(first row, rows available) of the example window (dataset.py:176-198) -- the integer part of WindowDataset._example."""
        L, W = self.example_window_length, self.window
        s0, e0 = int(self.ranges[ri][0]), int(self.ranges[ri][1])
        first, last = start, start + W - 1
        ext = (L - W) // 2
        ws, we = min(ext, first - s0), min(ext, e0 - last)
        s_ext, w_ext = ws + ext - we, we + ext - ws
        a = max(first - s_ext, s0)
        b = min(min(last + w_ext, e0) + 1, self.n_frames)
        return a, b - a

    def sample_batch(self, batchsize, device=None):
        from . import _lib
        idx = self.rs.randint(0, len(self.starts), size=batchsize)
        starts = self.starts[idx].astype(np.int32)
        T = self.window
        dev = self.device
        out = {n: torch.empty((batchsize, T) + shp, dtype=torch.float32, device=dev) for n, shp in zip(self.names, self.shapes)}
        a = _lib.GatherArgs(B=batchsize, T=T, n_arrays=len(self.names))
        for k, n in enumerate(self.names):
            a.src[k], a.dst[k], a.width[k] = self.dev_arrays[k].data_ptr(), out[n].data_ptr(), self.widths[k]
        keep = [torch.from_numpy(starts).to(dev)]
        a.start = keep[0].data_ptr()
        if self.style_encoding_type == "label":
            lab = torch.zeros(batchsize, self.nlabels)
            lab[torch.arange(batchsize), torch.as_tensor(self.labels[self.rng_idx[idx]]).long()] = 1.0
            out["style"] = lab.to(dev)
        else:
            L = self.example_window_length
            rng = [self._example_range(int(self.starts[i]), int(self.rng_idx[i])) for i in idx]
            for (_, n) in rng:
                if 2 * n < L:
                    raise _lib.ZeggsError("style example shorter than half the example window (dataset.py:201-203 cannot pad it)")
            ex = torch.empty((batchsize, L, self.ex_width), dtype=torch.float32, device=dev)
            keep += [torch.tensor([r[0] for r in rng], dtype=torch.int32).to(dev), torch.tensor([r[1] for r in rng], dtype=torch.int32).to(dev)]
            a.ex_out, a.L, a.ex_width, a.n_ex = ex.data_ptr(), L, self.ex_width, len(self._EX)
            for k, n in enumerate(self._EX):
                a.ex_src[k] = self.names.index(n)
            a.ex_start, a.ex_n = keep[1].data_ptr(), keep[2].data_ptr()
            out["style"] = ex
        _lib.check(_lib.lib().zeggs_window_gather(a, _lib.stream_ptr()), "zeggs_window_gather")
        return out


class DevicePrefetcher:
    """Double-buffered host->device input pipeline: `upload` enqueues the copies of the NEXT step's batch on a side stream
    so they overlap the current step's kernels; `acquire` makes the compute stream wait for them.  (The reference copies
    its 11 batch tensors synchronously at the top of every iteration, train.py:215-225.)  The device buffers are allocated
    once (grow-only, per tensor name) and re-used, so the steady state performs no allocation; a buffer is overwritten
    only after the step that read it has been enqueued two acquires ago (event-ordered, no host synchronisation)."""

    def __init__(self, device, nbuf=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.flat = [dict() for _ in range(nbuf)]       # name -> 1-D device buffer (capacity)
        self.free_ev = [None] * nbuf                     # recorded on the compute stream when the buffer may be overwritten
        self.n = 0
        self.last = None

    def upload(self, host_batch):
        k = self.n % len(self.flat)
        self.n += 1
        views = {}
        for name, v in host_batch.items():
            buf = self.flat[k].get(name)
            if buf is None or buf.numel() < v.numel() or buf.dtype != v.dtype:
                buf = torch.empty(max(v.numel(), 1), dtype=v.dtype, device=self.device)
                self.flat[k][name] = buf
            views[name] = buf[:v.numel()].view(v.shape)
        with torch.cuda.stream(self.stream):
            if self.free_ev[k] is not None:
                self.stream.wait_event(self.free_ev[k])
            for name, v in host_batch.items():
                views[name].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return k, views, ev

    def acquire(self, token):
        k, views, ev = token
        cur = torch.cuda.current_stream(self.device)
        if self.last is not None:                        # everything that read the previous buffer is enqueued by now
            fe = torch.cuda.Event()
            fe.record(cur)
            self.free_ev[self.last] = fe
        cur.wait_event(ev)
        self.last = k
        return views


class LaggedScalarReader:
    """Per-step device->host read of a scalar result (the loss) without draining the GPU: `push` enqueues a 4-byte copy into a
    pinned ring behind the step that produced the value and returns the values of the steps whose copies have certainly landed
    (those pushed `lag` calls ago); `drain` waits for the rest.  Every step's loss still reaches the host, one step late, so
    the next step's launches are enqueued while the current one runs (train.py:424-431 reads `loss.item()` synchronously)."""

    def __init__(self, device, lag=1, depth=8):
        self.device = torch.device(device)
        self.lag = lag
        self.ring = torch.empty(depth, dtype=torch.float32).pin_memory()
        self.pending = []                                 # (slot, event)
        self.n = 0

    def push(self, value):
        slot = self.n % self.ring.numel()
        self.n += 1
        self.ring[slot:slot + 1].copy_(value.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.pending.append((slot, ev))
        out = []
        while len(self.pending) > self.lag:
            s, e = self.pending.pop(0)
            e.synchronize()
            out.append(float(self.ring[s]))
        return out

    def drain(self):
        out = []
        while self.pending:
            s, e = self.pending.pop(0)
            e.synchronize()
            out.append(float(self.ring[s]))
        return out

