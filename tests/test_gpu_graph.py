"""the per-frame sequence ROI Align -> encoder -> association (tracking.py:304-329) captured ONCE
as a CUDA graph and replayed per frame.  The library's calls are plain launches on the caller's stream (no allocation, no
synchronisation on the small-launch paths), so torch.cuda.graph can record them together with an encoder's kernels."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import native, tracker_ref

pytestmark = pytest.mark.gpu

import alufe_b200  # noqa: E402,F401
from alufe_b200 import MultiStreamTracker, SHIPPED_CONF, roi, synth  # noqa: E402


def test_frame_sequence_replays_as_cuda_graph():
    Hf, Wf, H_in, W_in, n = synth.CONFIGS["c1"][:5]
    C, MD = 64, 16
    dev = torch.device("cuda", 0)
    ms = MultiStreamTracker(1, SHIPPED_CONF, max_tracks=64, max_dets=MD, device=dev)
    ref = tracker_ref.TrackerRef(SHIPPED_CONF)
    scene = synth.Scene(3, n, H_in, W_in, drop=0.1)
    feat_np = synth.feature_map(1, 1, C, Hf, Wf)
    # static buffers the graph reads; each frame's inputs are copied into them before the replay
    feat = torch.from_numpy(feat_np).to(dev)
    det = torch.zeros((MD, 6), device=dev)                                    # the detector's rows [x1,y1,x2,y2,conf,cls]
    n_det = torch.zeros(1, dtype=torch.int32, device=dev)
    boxes = torch.zeros((1, MD, 4), dtype=torch.float64, device=dev)
    confs = torch.zeros((1, MD), dtype=torch.float64, device=dev)
    embs = torch.zeros((1, MD, 128), dtype=torch.float32, device=dev)
    frame = torch.zeros(1, dtype=torch.int32, device=dev)
    result = torch.zeros((1, ms.stride), dtype=torch.int32, device=dev)
    proj = torch.randn((C * 100, 16), device=dev)                             # stand-in for the encoder (card.py:24-41)

    def frame_sequence():
        patches = roi.roi_align_from_input_boxes(feat, det, (H_in, W_in), out_size=(10, 10))
        side = patches.reshape(MD, -1) @ proj                                 # "encoder": consumes the patches on the same stream
        ms.step_device(n_det, boxes, confs, embs, frame, result)
        return patches, side

    side_stream = torch.cuda.Stream(dev)                                      # torch's capture protocol: warm up off the default stream
    side_stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side_stream):
        n_det.fill_(-1)                                                       # idle stream: the warm-up does not touch the tracker state
        frame_sequence()
    torch.cuda.current_stream(dev).wait_stream(side_stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        patches, side = frame_sequence()
    for f in range(12):
        obj = scene.step()
        k = len(obj["bboxes"])
        d = np.zeros((MD, 6), np.float32)
        d[:k, :4] = obj["bboxes"]
        det.copy_(torch.from_numpy(d))
        n_det.fill_(k)
        b = np.zeros((1, MD, 4)); b[0, :k] = obj["bboxes"]
        c = np.zeros((1, MD)); c[0, :k] = obj["confs"]
        e = np.zeros((1, MD, 128), np.float32); e[0, :k] = np.stack(obj["embs"])
        boxes.copy_(torch.from_numpy(b)); confs.copy_(torch.from_numpy(c)); embs.copy_(torch.from_numpy(e))
        frame.fill_(f)
        graph.replay()
        torch.cuda.synchronize()
        want = ref.update(obj)
        got = ms.decode(result[0].cpu().numpy())
        assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2], f
        rois = np.concatenate([np.zeros((MD, 1), np.float32), d[:, :4]], 1)
        assert_close(patches[:k].cpu().numpy(), native.roi_align(feat_np, rois[:k], (10, 10), Hf / float(H_in), 2, True),
                     what="graph replay frame %d" % f)
        assert torch.isfinite(side).all()
