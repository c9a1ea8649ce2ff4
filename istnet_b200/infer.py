"""Eval-mode inference path: what `test_func` (utils/solver.py:217-241) does per image — eval forward of IST_Net on
the B instances of one image (ist_net.py:67-74: no camera / world enhancers), then the 4x4 pose assembly
  scale = |pred_size|;  pred_scales = pred_size / scale;  pred_RTs = [[R * scale, t], [0, 1]]
— as ONE CUDA-graph launch per instance-count bucket: instance counts vary per image (1-10 objects), graphs need static shapes, so the
batch is padded to the next bucket size (instances are independent in eval mode: BatchNorm uses running statistics, so padding rows do
not influence the real ones) and the pose assembly runs on the device inside the graph.

BatchNorm needs no weight folding here: the eval BatchNorm is already fused into the single activation pass that follows every
convolution (nhwc.bn_act_split with running statistics), i.e. it costs no extra pass or launch.
"""
import torch

BUCKETS = (1, 2, 4, 8, 16, 32)


def assemble_poses(ep):
    """utils/solver.py:231-241 on the device: returns (pred_RTs [B,4,4], pred_scales [B,3])."""
    t, s, R = ep["pred_translation"], ep["pred_size"], ep["pred_rotation"]
    scale = torch.linalg.vector_norm(s, dim=1, keepdim=True)
    rts = torch.zeros(R.shape[0], 4, 4, dtype=R.dtype, device=R.device)
    rts[:, 3, 3] = 1.0
    rts[:, :3, 3] = t
    rts[:, :3, :3] = R * scale.unsqueeze(2)
    return rts, s / scale


class InferenceEngine:
    """engine = InferenceEngine(model, npts=1024, img=192);  pred_RTs, pred_scales = engine(inputs)  with inputs as test_func builds them
    ('rgb' [B,3,H,W], 'pts' [B,N,3], 'choose' [B,N] int64, 'category_label' [B] or [B,1] int64), any B <= max(BUCKETS)."""

    def __init__(self, model, npts=1024, img=192, device=None, use_graphs=True):
        self.model = model.eval()
        self.dev = device or next(model.parameters()).device
        self.npts, self.img, self.use_graphs = npts, img, use_graphs
        self._graphs = {}

    def _bucket(self, b):
        for s in BUCKETS:
            if b <= s:
                return s
        raise ValueError(f"more than {BUCKETS[-1]} instances in one image")

    def _run(self, inp):
        with torch.no_grad():
            return assemble_poses(self.model(inp))

    def _capture(self, size):
        st = {
            "rgb": torch.zeros(size, 3, self.img, self.img, device=self.dev),
            "pts": torch.zeros(size, self.npts, 3, device=self.dev),
            "choose": torch.zeros(size, self.npts, dtype=torch.int64, device=self.dev),
            "category_label": torch.zeros(size, 1, dtype=torch.int64, device=self.dev),
        }
        st["pts"][:, :, 2] = 1.0  # any finite cloud: the capture pass only records the launches
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._run(st)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self._run(st)
        return g, st, out

    def __call__(self, inputs):
        b = inputs["pts"].shape[0]
        if not self.use_graphs:
            return self._run({k: inputs[k] for k in ("rgb", "pts", "choose", "category_label")})
        size = self._bucket(b)
        if size not in self._graphs:
            self._graphs[size] = self._capture(size)
        g, st, out = self._graphs[size]
        st["rgb"][:b].copy_(inputs["rgb"], non_blocking=True)
        st["pts"][:b].copy_(inputs["pts"], non_blocking=True)
        st["choose"][:b].copy_(inputs["choose"], non_blocking=True)
        st["category_label"][:b].copy_(inputs["category_label"].reshape(b, 1), non_blocking=True)
        if b < size:  # padding rows repeat the first instance (finite values; eval-mode rows are independent)
            for k in st:
                st[k][b:size].copy_(st[k][0:1].expand_as(st[k][b:size]))
        g.replay()
        return out[0][:b], out[1][:b]


def estimate_poses(engine, rgb_frame, depth, valid_mask, det_boxes, class_ids, intrinsics, generator=None, norm_scale=1000.0):
    """One image of `test_func` (utils/solver.py:217-241) from DEVICE-resident frame data, without a host round trip: the per-instance
    preparation of the reference's test Dataset (provider/dataset.py:369-409; istnet_b200.dataprep: crop + 8-bit bilinear resize +
    normalise, valid-pixel sampling, back-projection, `choose` re-mapping) feeds the bucketed eval graph.
    rgb_frame [H,W,3] uint8 (RGB), depth [H,W] float32 (hole-filled), valid_mask [K,H,W] bool per detection (mask & depth > 0),
    det_boxes [K,4] host ints (y1,x1,y2,x2), class_ids [K] 0-indexed.  Instances with <= 16 valid pixels are dropped like the
    reference does (dataset.py:381).  Returns (pred_RTs [k,4,4], pred_scales [k,3], kept indices)."""
    from . import dataprep

    H, W = depth.shape
    keep = [j for j in range(len(det_boxes)) if int(valid_mask[j].sum()) > 16]
    if not keep:
        z = torch.zeros(0, device=depth.device)
        return z.view(0, 4, 4), z.view(0, 3), keep
    boxes_h = [(0,) + dataprep.get_bbox(det_boxes[j], H, W) for j in keep]
    choose = torch.empty(len(keep), engine.npts, dtype=torch.int32, device=depth.device)
    for i, j in enumerate(keep):  # each detection samples inside its own mask (dataset.py:375-386)
        c, ok = dataprep.sample_choose(valid_mask[j : j + 1], [boxes_h[i]], engine.npts, generator=generator)
        choose[i] = c[0]
    boxes = torch.tensor(boxes_h, dtype=torch.int32, device=depth.device)
    inp = dataprep.prepare_instances(rgb_frame.unsqueeze(0), depth.unsqueeze(0), boxes, choose, intrinsics, img_size=engine.img, norm_scale=norm_scale)
    inp["category_label"] = torch.as_tensor([int(class_ids[j]) for j in keep], dtype=torch.int64, device=depth.device)
    rts, scales = engine(inp)
    return rts, scales, keep
