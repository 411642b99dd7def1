"""GPU detection post-processing: the step right after `Model.predict_spectrogram` in the reference's
evaluation harness (microwakeword/test.py).  Same function names / arguments as the reference where it has
a function; the arithmetic runs in libmww_b200.so (csrc/mww_detect.cu) so that only detection counts and
scores -- not every probability -- have to leave the GPU.

    compute_false_accepts_per_hour(tracks, cutoffs, ignore_slices_after_accept=75, stride=1, step_s=0.02)
        == microwakeword.test.compute_false_accepts_per_hour (test.py:94-137) applied to the moving averages
           of `tracks` (pass window=1 to feed already-averaged tracks like the reference call site does)
    moving_average(tracks, window=5)          test.py:337-341
    positive_scores(tracks, window=5, ignore) test.py:364-373
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _pack(tracks, device):
    import torch
    lengths = np.asarray([len(t) for t in tracks], np.int32)
    offsets = np.zeros(len(tracks), np.int64)
    if len(tracks) > 1:
        offsets[1:] = np.cumsum(lengths[:-1], dtype=np.int64)
    if len(tracks) and all(isinstance(t, torch.Tensor) for t in tracks):
        flat = torch.cat([t.to(device=device, dtype=torch.float32).reshape(-1) for t in tracks]) if len(tracks) else torch.zeros(0, device=device)
    else:
        flat = torch.from_numpy(np.concatenate([np.asarray(t, np.float32).reshape(-1) for t in tracks]) if len(tracks) else np.zeros(0, np.float32)).to(device)
    return flat, torch.from_numpy(offsets).to(device), torch.from_numpy(lengths).to(device), lengths


def _stream(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def moving_average(tracks, window: int = 5, device: int = 0):
    """list of 1-D probability tracks -> list of float32 numpy arrays (len - window + 1 each)."""
    import torch
    dev = torch.device("cuda", device)
    flat, offs, lens_d, lens = _pack(tracks, dev)
    out_len = np.maximum(lens.astype(np.int64) - window + 1, 0)
    out_off = np.zeros(len(tracks), np.int64)
    if len(tracks) > 1:
        out_off[1:] = np.cumsum(out_len[:-1])
    out = torch.empty(max(int(out_len.sum()), 1), dtype=torch.float32, device=dev)
    rc = _lib.lib().mww_moving_average(flat.data_ptr(), offs.data_ptr(), lens_d.data_ptr(), len(tracks), int(lens.max()) if len(tracks) else 0, window,
                                       out.data_ptr(), torch.from_numpy(out_off).to(dev).data_ptr(), _stream(dev))
    if rc != 0:
        raise _lib.MwwError(rc, "mww_moving_average failed")
    host = out.cpu().numpy()
    return [host[o:o + n].copy() for o, n in zip(out_off, out_len)]


def false_accept_counts(tracks, cutoffs, ignore_slices_after_accept: int = 75, window: int = 1, device: int = 0) -> np.ndarray:
    """int32 [n_tracks, n_cutoffs] detections (cooldown rule of test.py:120-135) on the `window`-point moving average."""
    import torch
    dev = torch.device("cuda", device)
    flat, offs, lens_d, _ = _pack(tracks, dev)
    cut = torch.from_numpy(np.ascontiguousarray(cutoffs, np.float64)).to(dev)
    counts = torch.zeros((max(len(tracks), 1), cut.numel()), dtype=torch.int32, device=dev)
    rc = _lib.lib().mww_false_accept_counts(flat.data_ptr(), offs.data_ptr(), lens_d.data_ptr(), len(tracks), window, cut.data_ptr(), cut.numel(),
                                            int(ignore_slices_after_accept), counts.data_ptr(), _stream(dev))
    if rc != 0:
        raise _lib.MwwError(rc, "mww_false_accept_counts failed")
    return counts.cpu().numpy()[:len(tracks)]


def compute_false_accepts_per_hour(streaming_probabilities_list, cutoffs, ignore_slices_after_accept: int = 75, stride: int = 1,
                                   step_s: float = 0.02, window: int = 1, device: int = 0) -> np.ndarray:
    """Drop-in for microwakeword.test.compute_false_accepts_per_hour (test.py:94-137).  With the default window=1 the
    tracks are used as given (the reference passes moving averages); window=5 fuses test.py:337-341 into the same pass,
    in which case the duration is counted on the averaged length like the reference does."""
    cutoffs = np.asarray(cutoffs, np.float64)
    counts = false_accept_counts(streaming_probabilities_list, cutoffs, ignore_slices_after_accept, window, device)
    duration_h = 0
    for t in streaming_probabilities_list:
        duration_h += max(len(t) - window + 1, 0) * stride * step_s / 3600.0
    return counts.sum(0).astype(np.float64) / duration_h


def positive_scores(tracks, window: int = 5, ignore_slices_after_accept: int = 25, device: int = 0) -> np.ndarray:
    """float32 [n_tracks]: max of the moving average after skipping the first `ignore` probabilities (test.py:364-373)."""
    import torch
    dev = torch.device("cuda", device)
    flat, offs, lens_d, _ = _pack(tracks, dev)
    out = torch.empty(max(len(tracks), 1), dtype=torch.float32, device=dev)
    rc = _lib.lib().mww_positive_scores(flat.data_ptr(), offs.data_ptr(), lens_d.data_ptr(), len(tracks), window, int(ignore_slices_after_accept),
                                        out.data_ptr(), _stream(dev))
    if rc != 0:
        raise _lib.MwwError(rc, "mww_positive_scores failed")
    return out.cpu().numpy()[:len(tracks)]
