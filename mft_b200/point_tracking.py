"""Point tracks from a dense result: the reference's MFT/point_tracking.py surface (point_tracking.py:6-27).

The reference samples the full field on whatever device the result lives on -- in its runners that is the CPU copy
returned by track(), i.e. a 4*H*W*4-byte device->host transfer per frame for a few hundred points.
Here a CUDA result (``track(..., device_result=True)``) is sampled by ONE launch of the library's point-sampling
kernel (flow + occlusion in the same pass) and only (N, 2) + (N,) floats cross to the host."""
import numpy as np
import torch

from .results import _as_f32


def convert_to_point_tracking(MFT_result, queries):
    """args:    MFT_result: FlowOUTrackingResult (template -> current frame); queries: (N, xy) positions in the template
    returns: current_coords (N, xy) float32 numpy, current_occlusions (N,) float32 numpy -- bilinear samples with
             zeros outside the image (grid_sample, align_corners=True), like the reference."""
    q = _as_f32(queries)
    s = MFT_result._sample_points(MFT_result.packed()[0:3], q, add_points=True)      # rows: x + fx, y + fy, occlusion
    s = s.detach().cpu().numpy() if isinstance(s, torch.Tensor) else np.asarray(s)
    return np.ascontiguousarray(s[0:2].T, dtype=np.float32), np.ascontiguousarray(s[2], dtype=np.float32)
