"""IST-Net top-level modules on the B200 kernels, behind the reference's module surface.

Same class names, constructor signatures, forward dict keys, train/eval behaviour and `state_dict` keys as
`model/ist_net.py` and `model/posenet_gt.py` of the reference, so `train.py` / `test.py` run unchanged when
`compat/` is on sys.path.  Differences from the reference are deliberate and documented:
no hard-coded `.cuda()` (everything follows the input's device), no weight download.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import rows_engine as RE
from . import trace
from .image import ModifiedResnet
from .pointnet2 import PointNet2MSG

CAM_RADII = [[0.01, 0.02], [0.02, 0.04], [0.04, 0.08], [0.08, 0.16]]  # ist_net.py:16, posenet_gt.py:18
WORLD_RADII = [[0.05, 0.10], [0.10, 0.20], [0.20, 0.30], [0.30, 0.40]]  # ist_net.py:189, posenet_gt.py:19


# --------------------------------------------------------------------------------------------- stream-level concurrency
USE_SIDE_STREAMS = os.environ.get("ISTNET_STREAMS", "1") != "0"
# "points": the two point-cloud streams (extractors, enhancer heads) outrank the image stream.  Their kernels are short and
# latency-bound; behind the image branch's persistent full-GPU kernels they starve (measured: the 5 ms camera-extractor
# backward stretched over 15.6 ms and ended 3 ms after the image backward, tools/timeline.py), ahead of them they cost the
# image branch little.  "image": the image stream outranks them (round-1 default until this measurement).  "none": equal.
PRIORITY_MODE = os.environ.get("ISTNET_PRIO", "image")
# pose heads: enqueue the main path (implicit transformation -> main estimator) before the two enhancer heads so that
# autograd (which replays nodes newest-first and makes a consumer stream wait for everything already enqueued on the
# producer stream) starts the enhancers' backward passes right after the loss instead of behind the main path's backward
HEADS_MAIN_FIRST = os.environ.get("ISTNET_HEADS_MAIN_FIRST", "1") != "0"
MAIN_PATH_HIGH_PRIORITY = os.environ.get("ISTNET_MAIN_HI", "1") == "1"


class _Branches:
    """Runs independent sub-networks (image branch, camera-space extractor, NOCS-space extractor) on side CUDA streams so
    that the latency-bound point-cloud kernels overlap the tensor-core-bound image branch.  Autograd replays each
    backward node on its forward stream, so the overlap carries over to the backward pass; under CUDA-graph capture the
    fork / join become graph dependencies."""

    def __init__(self, device, n):
        self.main = torch.cuda.current_stream(device)
        self.on = USE_SIDE_STREAMS and device.type == "cuda"
        self.side = []
        if self.on:
            pool = _Branches._pool.setdefault(device, [])
            while len(pool) < n:  # lower number = higher priority; PRIORITY_MODE picks who wins when CTAs compete for an SM
                i = len(pool)
                hi = (i == 2) if PRIORITY_MODE == "image" else (i < 2) if PRIORITY_MODE == "points" else False
                pool.append(torch.cuda.Stream(device, priority=-1 if hi else 0))
            self.side = pool[:n]
            for st in self.side:
                st.wait_stream(self.main)
        self.results = []

    _pool = {}

    def run(self, i, fn):
        if not self.on:
            return fn()
        with torch.cuda.stream(self.side[i]):
            out = fn()
        self.results.append((self.side[i], out))
        return out

    def join(self):
        for st, out in self.results:
            self.main.wait_stream(st)
            for t in out if isinstance(out, (tuple, list)) else (out,):
                if isinstance(t, torch.Tensor):
                    t.record_stream(self.main)


def check_fp32_matmul():
    """The remaining library matmuls on the path (PSP prior maps, nn.Linear pose heads; < 0.1 % of the FLOPs) must run in true
    FP32 for the 1e-4 parity target.  That is PyTorch's default; importing this package does not touch the global flag, it only
    refuses to run with it switched on."""
    if torch.backends.cuda.matmul.allow_tf32:
        raise RuntimeError("istnet_b200: torch.backends.cuda.matmul.allow_tf32 is True; the 1e-4 parity target of the pose heads needs "
                           "FP32 matmuls (PyTorch's default) — set it to False around the model's forward/backward")


# --------------------------------------------------------------------------------------------- small pieces
def ortho6d_to_mat(x_raw, y_raw):
    """Ortho6d2Mat (utils/rotation_utils.py:4-28): y=norm(y_raw); z=norm(x_raw x y); x=y x z; columns [x,y,z].
    Same arithmetic as the reference (sqrt of the sum of squares clamped at 1e-8, component-wise cross products) in 9
    launches instead of 31: the function runs four times per step and its ~100 forward+backward micro-kernels each cost
    a graph-node latency on the pose heads' critical path."""

    def nrm(v):
        return v / torch.linalg.vector_norm(v, dim=1, keepdim=True).clamp_min(1e-8)

    y = nrm(y_raw)
    z = nrm(torch.linalg.cross(x_raw, y, dim=1))
    x = torch.linalg.cross(y, z, dim=1)
    return torch.stack((x, y, z), 2)


def _pt_mlp(*widths, last_relu=True):
    """[Conv1d(k=1)+bias, ReLU]* as an nn.Sequential with the reference's child indices (0,2,4,...)."""
    layers = []
    for i in range(len(widths) - 1):
        layers.append(nn.Conv1d(widths[i], widths[i + 1], 1))
        if last_relu or i + 2 < len(widths):
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


def _head(out):
    return nn.Sequential(nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, out))


def gather_pixels(rgb_map, choose):
    """ist_net.py:42-45: (B,d,H,W), choose (B,N) int64 -> (B,d,N)"""
    b, d = rgb_map.shape[:2]
    return torch.gather(rgb_map.view(b, d, -1), 2, choose.unsqueeze(1).expand(-1, d, -1)).contiguous()


def _rows(x_bnc):
    """(B,N,C) -> contiguous row matrix (B*N, C)"""
    b, n, c = x_bnc.shape
    return x_bnc.reshape(b * n, c)


# Operand planes of the pose heads' per-point MLPs (BatchNorm-free Conv1d+ReLU stacks, <= 10 layers deep): see DESIGN.md section 2
HEADS_NSPLIT = int(os.environ.get("ISTNET_NSPLIT_HEADS", "2"))


def _mlp(seq, x_rows, training=False):
    """nn.Sequential of Conv1d(k=1)(+ReLU) applied to a row matrix on the tcgen05 GEMM chain (rows_engine)."""
    return RE.run_chain(RE.units_from_conv1d_seq(seq, nsplit=min(HEADS_NSPLIT, RE.K.NSPLIT)), x_rows, training)


def _with_global(x_rows, b, n):
    """[f, mean_over_points(f).expand] of ist_net.py:172-173,257-258,325-326 on rows (the concatenation happens in the operand split)"""
    f = x_rows.view(b, n, -1)
    return [x_rows, f.mean(1, keepdim=True).expand_as(f).reshape(b * n, -1)]


# The estimators' global feature as a per-instance bias (nhwc.ConvUnit.glob): W [f | mean(f)] = W_a f + (W_b mean(f) + b) — halves the K of
# the first layer of deform_mlp2 / pose_mlp2 (forward, data gradient, weight gradient) and removes the expanded [rows, 256] tensor
GLOBAL_BIAS = os.environ.get("ISTNET_GLOBAL_BIAS", "1") != "0"


def _mlp_global(seq, x_rows, b, n):
    """seq(torch.cat([f, mean_over_points(f).expand], channels)) of ist_net.py:172-173,257-258,325-326 on rows."""
    if GLOBAL_BIAS and x_rows.is_cuda and b <= 64 and x_rows.shape[1] % 8 == 0:
        units = RE.units_from_conv1d_seq(seq, nsplit=min(HEADS_NSPLIT, RE.K.NSPLIT))
        units[0].glob = n
        return RE.run_chain(units, x_rows, False)
    return _mlp(seq, _with_global(x_rows, b, n))


FUSED_HEADS = os.environ.get("ISTNET_FUSED_HEADS", "1") != "0"


def _ptr_array(tensors):
    import ctypes

    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])


class _PoseTailFn(torch.autograd.Function):
    """AdaptiveAvgPool1d(1) -> three [Linear, ReLU, Linear, ReLU, Linear] heads -> Ortho6d2Mat on the kernels of csrc/heads.cu:
    5 launches forward, 6 backward for all three heads (ist_net.py:228-264,296-332; ~25 / ~70 library launches in the reference)."""

    @staticmethod
    def forward(ctx, feat_rows, b, n, *params):
        import ctypes

        from . import _C
        from ._C import c_int, ptr

        dev = feat_rows.device
        C = feat_rows.shape[1]
        pooled = torch.empty(b, C, dtype=torch.float32, device=dev)
        _C.call("rows_mean", c_int(b), c_int(n), c_int(C), ptr(feat_rows), ptr(pooled))
        W = [[params[6 * h + 2 * l] for l in range(3)] for h in range(3)]
        Bs = [[params[6 * h + 2 * l + 1] for l in range(3)] for h in range(3)]
        xs, ys = [pooled] * 3, []
        for l in range(3):
            O = [W[h][l].shape[0] for h in range(3)]
            y = [torch.empty(b, O[h], dtype=torch.float32, device=dev) for h in range(3)]
            _C.call("heads_linear", c_int(3), c_int(b), c_int(W[0][l].shape[1]), _ptr_array(xs), _ptr_array([W[h][l] for h in range(3)]),
                    _ptr_array([Bs[h][l] for h in range(3)]), _ptr_array(y), (ctypes.c_int * 3)(*O), c_int(1 if l < 2 else 0))
            ys.append(y)
            xs = y
        R = torch.empty(b, 3, 3, dtype=torch.float32, device=dev)
        _C.call("ortho6d", c_int(b), ptr(ys[2][0]), ptr(R))
        ctx.saved = (pooled, ys, W, Bs)
        ctx.dims = (b, n, C)
        ctx.params = params
        return R, ys[2][1], ys[2][2]

    @staticmethod
    def backward(ctx, dR, dt, ds):
        import ctypes

        from . import _C
        from ._C import c_int, c_ll, ptr

        pooled, ys, W, Bs = ctx.saved
        b, n, C = ctx.dims
        dev = pooled.device
        f32 = dict(dtype=torch.float32, device=dev)
        dR = torch.zeros(b, 3, 3, **f32) if dR is None else dR.contiguous()
        dt = torch.zeros(b, 3, **f32) if dt is None else dt.contiguous()
        ds = torch.zeros(b, 3, **f32) if ds is None else ds.contiguous()
        dr6 = torch.empty(b, 6, **f32)
        _C.call("ortho6d_bwd", c_int(b), ptr(ys[2][0]), ptr(dR), ptr(dr6))
        dy = [dr6, dt, ds]
        gW = [[None] * 3 for _ in range(3)]
        gB = [[None] * 3 for _ in range(3)]
        for l in (2, 1, 0):
            x = ys[l - 1] if l > 0 else [pooled] * 3
            K_ = W[0][l].shape[1]
            O = [W[h][l].shape[0] for h in range(3)]
            dx = [torch.empty(b, K_, **f32) for _ in range(3)]
            for h in range(3):
                gW[h][l], gB[h][l] = torch.empty_like(W[h][l]), torch.empty_like(Bs[h][l])
            _C.call("heads_linear_bwd", c_int(3), c_int(b), c_int(K_), _ptr_array(x), _ptr_array([W[h][l] for h in range(3)]),
                    _ptr_array(ys[l]), _ptr_array(dy), _ptr_array(dx), _ptr_array([gW[h][l] for h in range(3)]),
                    _ptr_array([gB[h][l] for h in range(3)]), (ctypes.c_int * 3)(*O), c_int(1 if l < 2 else 0))
            dy = dx
        dpooled = torch.empty(b, C, **f32)
        _C.call("sum3", c_ll(b * C), ptr(dy[0]), ptr(dy[1]), ptr(dy[2]), ptr(dpooled))
        dfeat = torch.empty(b * n, C, **f32)
        _C.call("rows_mean_bwd", c_int(b), c_int(n), c_int(C), ptr(dpooled), ptr(dfeat))
        grads = []
        for h in range(3):
            for l in range(3):
                grads += [gW[h][l], gB[h][l]]
        ctx.saved = None
        return (dfeat, None, None) + tuple(g if p.requires_grad else None for g, p in zip(grads, ctx.params))


class _PoseHeads(nn.Module):
    """pose_mlp1 -> global mean concat -> pose_mlp2 -> avg pool -> rotation / translation / size heads."""

    def _head_params(self):
        ps = []
        for head in (self.rotation_estimator, self.translation_estimator, self.size_estimator):
            for i in (0, 2, 4):
                ps += [head[i].weight, head[i].bias]
        return ps

    def _tail(self, feat_rows, b, n):
        feat = _mlp(self.pose_mlp1, feat_rows)
        feat = _mlp_global(self.pose_mlp2, feat, b, n)
        if FUSED_HEADS and feat.is_cuda and b <= 64 and feat.shape[1] % 4 == 0:
            feat = feat.contiguous()
            params = self._head_params()
            if torch.is_grad_enabled() and (feat.requires_grad or any(p.requires_grad for p in params)):
                return _PoseTailFn.apply(feat, b, n, *params)
            return _PoseTailFn.forward(RE._NoCtx(), feat, b, n, *params)
        feat = feat.view(b, n, -1).mean(1)  # AdaptiveAvgPool1d(1)
        r6 = self.rotation_estimator(feat)
        r = ortho6d_to_mat(r6[:, :3].contiguous(), r6[:, 3:].contiguous()).view(-1, 3, 3)
        return r, self.translation_estimator(feat), self.size_estimator(feat)


class LightEstimator(_PoseHeads):
    """ist_net.py:202-264.  Inputs are rows: pts (B,N,3), rgb_local / pts_local (B,N,128)."""

    def __init__(self):
        super().__init__()
        self.pts_mlp = _pt_mlp(3, 32, 64)
        self.pose_mlp1 = _pt_mlp(128 + 64 + 128, 256, 256)
        self.pose_mlp2 = nn.Sequential(nn.Conv1d(512, 512, 1), nn.ReLU(), nn.Conv1d(512, 512, 1), nn.ReLU(), nn.AdaptiveAvgPool1d(1))
        self.rotation_estimator, self.translation_estimator, self.size_estimator = _head(6), _head(3), _head(3)

    def forward(self, pts, rgb_local, pts_local):
        b, n, _ = pts.shape
        e = _mlp(self.pts_mlp, _rows(pts))
        return self._tail([_rows(rgb_local), e, _rows(pts_local)], b, n)


class HeavyEstimator(_PoseHeads):
    """ist_net.py:267-332 (identical copy at posenet_gt.py:71-136).  Inputs are rows (B,N,C)."""

    def __init__(self):
        super().__init__()
        self.pts_mlp1 = _pt_mlp(3, 32, 64)
        self.pts_mlp2 = _pt_mlp(3, 32, 64)
        self.pose_mlp1 = _pt_mlp(64 + 64 + 384, 256, 256)
        self.pose_mlp2 = nn.Sequential(nn.Conv1d(512, 512, 1), nn.ReLU(), nn.Conv1d(512, 512, 1), nn.ReLU(), nn.AdaptiveAvgPool1d(1))
        self.rotation_estimator, self.translation_estimator, self.size_estimator = _head(6), _head(3), _head(3)

    def forward(self, pts, pts_w, rgb_local, pts_local, pts_w_local):
        b, n, _ = pts.shape
        e1 = _mlp(self.pts_mlp1, _rows(pts))
        e2 = _mlp(self.pts_mlp2, _rows(pts_w))
        return self._tail([_rows(rgb_local), e1, _rows(pts_local), e2, _rows(pts_w_local)], b, n)


class FeatureDeformer(nn.Module):
    """Implicit space transformation (ist_net.py:125-183).  Inputs are rows; returns (pts_w_local rows (B,N,128), pts_w (B,N,3))."""

    def __init__(self, nclass=6):
        super().__init__()
        self.nclass = nclass
        self.pts_mlp1 = _pt_mlp(3, 32, 64)
        self.deform_mlp1 = _pt_mlp(64 + 256, 384, 256)
        self.deform_mlp2 = _pt_mlp(512, 384, 256, 128)
        self.pred_nocs = _pt_mlp(128, 256, 128, nclass * 3, last_relu=False)

    def forward(self, pts, rgb_local, pts_local, cls):
        b, n, _ = pts.shape
        e = _mlp(self.pts_mlp1, _rows(pts))
        x = _mlp(self.deform_mlp1, [e, _rows(pts_local), _rows(rgb_local)])
        x = _mlp_global(self.deform_mlp2, x, b, n)
        q = _mlp(self.pred_nocs, x).view(b, n, self.nclass, 3)
        # ist_net.py:178-181: view(-1,3,N) + index_select(cls + nclass*b) == pick the class's 3 channels per instance
        q = torch.gather(q, 2, cls.view(b, 1, 1, 1).expand(b, n, 1, 3)).squeeze(2)
        return x.view(b, n, -1), q.contiguous()


class ImplicitTransformation(nn.Module):
    """ist_net.py:114-122"""

    def __init__(self, nclass=6):
        super().__init__()
        self.nclass = nclass
        self.feature_refine = FeatureDeformer(nclass)

    def forward(self, rgb_local, pts_local, pts, center, cls):
        pts_local_w, pts_w = self.feature_refine(pts, rgb_local, pts_local, cls)
        return pts_w, pts_local_w


class WorldSpaceEnhancer(nn.Module):
    """ist_net.py:185-200 (rows in / rows out)"""

    def __init__(self, freeze=False):
        super().__init__()
        self.freeze = freeze
        self.extractor = PointNet2MSG(radii_list=WORLD_RADII)
        if not freeze:
            self.pose_estimator = HeavyEstimator()

    def forward(self, pts, pts_w_gt, rgb_local, pts_local, pts_w_local_gt=None):
        if pts_w_local_gt is None:  # (IST_Net runs the extractor on a side stream and passes its output in)
            pts_w_local_gt = self.extractor.forward_rows(pts_w_gt)
        if self.freeze:
            return None, None, None, pts_w_local_gt
        r, t, s = self.pose_estimator(pts, pts_w_gt, rgb_local.detach(), pts_local.detach(), pts_w_local_gt)
        return r, t, s, pts_w_local_gt


# --------------------------------------------------------------------------------------------- top modules
class IST_Net(nn.Module):
    """ist_net.py:10-76"""

    def __init__(self, nclass=6, freeze_world_enhancer=False):
        super().__init__()
        self.nclass = nclass
        self.freeze_world_enhancer = freeze_world_enhancer
        self.rgb_cam_extractor = ModifiedResnet()
        self.pts_cam_extractor = PointNet2MSG(radii_list=CAM_RADII)
        self.implicit_transform = ImplicitTransformation(nclass)
        self.main_estimator = HeavyEstimator()
        self.cam_enhancer = LightEstimator()
        self.world_enhancer = WorldSpaceEnhancer(freeze=freeze_world_enhancer)

    def forward(self, inputs):
        end_points = {}
        check_fp32_matmul()
        RE.K.begin_step(self, inputs["pts"].device)  # all weights of the model re-laid as operand planes in one launch (nhwc.WeightBank)
        rgb, pts, choose = inputs["rgb"], inputs["pts"], inputs["choose"]
        cls = inputs["category_label"].reshape(-1)
        c = torch.mean(pts, 1, keepdim=True)
        pts = pts - c
        # everything below works on rows (B,N,C): the layout the GEMM kernels consume; the reference's (B,C,N)
        # tensors appear only at the module boundary (end_points)
        br = _Branches(pts.device, 3)
        ph = trace.phase
        rgb_local = br.run(2, lambda: ph("image", lambda: self.rgb_cam_extractor.gather_rows(rgb, choose)))
        pts_local = br.run(0, lambda: ph("cam_extractor", lambda: self.pts_cam_extractor.forward_rows(pts)))
        if self.training:  # the NOCS-space extractor only depends on the ground-truth coordinates
            gt_feats = br.run(1, lambda: ph("world_extractor", lambda: self.world_enhancer.extractor.forward_rows(inputs["qo"])))
        br.join()
        # the three pose heads are independent of each other: camera-space enhancer and world-space enhancer on the side
        # streams, implicit space transformation -> main estimator on the main stream
        br2 = _Branches(pts.device, 3)  # the side streams' wait on the main stream is taken here, before the main path is enqueued

        def _main_path():
            pw, pwl = ph("implicit_transform", lambda: self.implicit_transform(rgb_local, pts_local, pts, c, cls))
            return (pw, pwl) + tuple(ph("main_estimator", lambda: self.main_estimator(pts, pw, rgb_local, pts_local, pwl)))

        # MAIN_PATH_HIGH_PRIORITY: the chain implicit transformation -> main estimator (forward) and its backward gate the image
        # branch's backward, the longest chain of the step; on the high-priority stream (idle between the image branch's forward
        # and backward) it is not slowed down by the two enhancer heads running beside it
        run_main = (lambda: br2.run(2, _main_path)) if (MAIN_PATH_HIGH_PRIORITY and self.training) else _main_path
        if HEADS_MAIN_FIRST:
            pts_w, pts_w_local, r, t, s = run_main()
        if self.training:
            r_c, t_c, s_c = br2.run(0, lambda: ph("cam_enhancer", lambda: self.cam_enhancer(pts, rgb_local, pts_local)))
            r_w, t_w, s_w, pts_w_local_gt = br2.run(1, lambda: ph("world_enhancer", lambda: self.world_enhancer(pts, inputs["qo"], rgb_local, pts_local, gt_feats)))
        if not HEADS_MAIN_FIRST:
            pts_w, pts_w_local, r, t, s = run_main()
        br2.join()
        trace.mark("forward joined")
        end_points["pred_qo"] = pts_w
        if self.training:
            end_points["pts_w_local"] = pts_w_local.transpose(1, 2).contiguous()
            end_points["pts_w_local_gt"] = pts_w_local_gt.transpose(1, 2).contiguous()
        end_points["pred_rotation"] = r
        end_points["pred_translation"] = t + c.squeeze(1)
        end_points["pred_size"] = s
        if self.training:
            end_points["pred_rotation_aux_cam"] = r_c
            end_points["pred_translation_aux_cam"] = t_c + c.squeeze(1)
            end_points["pred_size_aux_cam"] = s_c
            if not self.freeze_world_enhancer:
                end_points["pred_rotation_aux_world"] = r_w
                end_points["pred_translation_aux_world"] = t_w + c.squeeze(1)
                end_points["pred_size_aux_world"] = s_w
        return end_points


class PoseNetGT(nn.Module):
    """posenet_gt.py:11-51 (phase-1 "world enhancer" training): only pts_gt_extractor and the estimator
    receive gradients; the rgb / camera-space extractors still run (BN running stats, dropout draws)."""

    def __init__(self, nclass=6, nprior=1024):
        super().__init__()
        self.nclass, self.nprior = nclass, nprior
        self.rgb_extractor = ModifiedResnet()
        self.pts_extractor = PointNet2MSG(radii_list=CAM_RADII)
        self.pts_gt_extractor = PointNet2MSG(radii_list=WORLD_RADII)
        self.pose_estimator_aux = HeavyEstimator()

    def forward(self, inputs):
        check_fp32_matmul()
        RE.K.begin_step(self, inputs["pts"].device)
        rgb, pts, choose, pts_w_gt = inputs["rgb"], inputs["pts"], inputs["choose"], inputs["qo"]
        c = torch.mean(pts, 1, keepdim=True)
        pts = pts - c
        br = _Branches(pts.device, 2)

        def _cam():
            with torch.no_grad():
                return self.pts_extractor.forward_rows(pts)

        pts_local = br.run(0, _cam)
        gt_local = br.run(1, lambda: self.pts_gt_extractor.forward_rows(pts_w_gt))
        with torch.no_grad():  # outputs are detached in the reference (posenet_gt.py:43); same values, no graph
            rgb_local = self.rgb_extractor.gather_rows(rgb, choose)
        br.join()
        r, t, s = self.pose_estimator_aux(pts, pts_w_gt, rgb_local, pts_local, gt_local)
        return {"pts_local_w_gt": gt_local.transpose(1, 2).contiguous(), "pred_rotation": r, "pred_translation": t + c.squeeze(1), "pred_size": s}


# --------------------------------------------------------------------------------------------- losses
def SmoothL1Dis(p1, p2, threshold=0.1):
    """losses.py:3-22"""
    diff = torch.abs(p1 - p2)
    dis = torch.where(diff > threshold, diff - threshold / 2.0, torch.pow(diff, 2) / (2.0 * threshold))
    return torch.mean(torch.sum(dis, dim=2 if p1.dim() == 3 else 1))


def PoseDis(r1, t1, s1, r2, t2, s2):
    """losses.py:37-49"""
    nrm = torch.linalg.vector_norm
    return torch.mean(nrm(r1 - r2, dim=1)) + torch.mean(nrm(t1 - t2, dim=1)) + torch.mean(nrm(s1 - s2, dim=1))


class SupervisedLoss(nn.Module):
    """ist_net.py:78-111.  `cfg` needs `.loss.gamma1`, `.loss.gamma2`, `.freeze_world_enhancer`."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg.loss
        self.freeze_world_enhancer = cfg.freeze_world_enhancer

    def forward(self, ep):
        R, T, S = ep["rotation_label"], ep["translation_label"], ep["size_label"]
        loss_feat = F.mse_loss(ep["pts_w_local"], ep["pts_w_local_gt"])
        loss_qo = SmoothL1Dis(ep["pred_qo"], ep["qo"])
        loss = PoseDis(ep["pred_rotation"], ep["pred_translation"], ep["pred_size"], R, T, S)
        loss = loss + PoseDis(ep["pred_rotation_aux_cam"], ep["pred_translation_aux_cam"], ep["pred_size_aux_cam"], R, T, S)
        loss = loss + self.cfg.gamma1 * loss_qo + self.cfg.gamma2 * loss_feat
        if not self.freeze_world_enhancer:
            loss = loss + PoseDis(ep["pred_rotation_aux_world"], ep["pred_translation_aux_world"], ep["pred_size_aux_world"], R, T, S)
        return loss


class PoseNetGTLoss(nn.Module):
    """posenet_gt.py:53-67"""

    def __init__(self, cfg=None):
        super().__init__()
        self.cfg = getattr(cfg, "loss", None)

    def forward(self, ep):
        return PoseDis(ep["pred_rotation"], ep["pred_translation"], ep["pred_size"], ep["rotation_label"], ep["translation_label"], ep["size_label"])


class LossCfg:
    """Minimal stand-in for the gorilla Config object (`cfg.loss.gamma1`, config/ist_net_default.yaml:25-30)."""

    class _L:
        def __init__(self, g1, g2):
            self.gamma1, self.gamma2 = g1, g2

    def __init__(self, gamma1=1.0, gamma2=10.0, freeze_world_enhancer=False):
        self.loss = LossCfg._L(gamma1, gamma2)
        self.freeze_world_enhancer = freeze_world_enhancer
