"""ORACLE — test infrastructure only (see oracle/__init__.py).  Never imported by istnet_b200/.

Functional, plain-PyTorch (FP32) restatement of the reference's per-instance hot path, driven directly by a
reference-layout `state_dict`.  It exists because the reference tree cannot travel to the GPU
box: it is (1) the float oracle for the `-m gpu` parity tests, (2) the `cpu_baseline` / `--impl reference`
arm of bench.py (kind "port").  It is pinned against the UNMODIFIED reference modules imported through
oracle/ref_harness.py (tests/test_oracle_model.py, runs where /root/reference exists) and against
the golden vectors in tests/golden/ that were produced by the reference itself.

Each function cites the reference lines it follows.  `ops` is a module exposing the nine `_ext` operators
(default: oracle.pointops, the C restatement).
"""
import torch
import torch.nn.functional as F

from . import pointops as _cpu_ops


class Ctx:
    """Run-time switches shared by all functions: train/eval, BN momentum source, dropout masks."""

    def __init__(self, sd, training, ops=None, bn_momentum=0.1, dropout_noise=None, update_stats=True, bn_training=None):
        self.sd = sd
        self.training = training
        self.ops = ops or _cpu_ops
        self.bn_momentum = bn_momentum
        self.dropout_noise = dropout_noise  # optional list of (B,C,1,1) tensors consumed in call order
        self._drop_i = 0
        self.update_stats = update_stats
        self.bn_training = training if bn_training is None else bn_training  # BatchNorm modules may be in eval() alone

    def p(self, key):
        return self.sd[key]


def batch_norm(cx, x, prefix):
    """nn.BatchNorm2d forward (eps 1e-5); train: batch stats + running update."""
    rm, rv = cx.p(prefix + ".running_mean"), cx.p(prefix + ".running_var")
    if cx.bn_training and not cx.update_stats:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, cx.p(prefix + ".weight"), cx.p(prefix + ".bias"), cx.bn_training, cx.bn_momentum, 1e-5)
    if cx.bn_training and cx.update_stats:
        cx.sd[prefix + ".num_batches_tracked"] += 1
    return y


def dropout2d(cx, x, p):
    """F.dropout2d == x * (bernoulli(1-p)/(1-p)) per (b,c) (modules.py:56,62,72-78)."""
    if not cx.training:
        return x
    if cx.dropout_noise is not None:
        noise = cx.dropout_noise[cx._drop_i]
        cx._drop_i += 1
        return x * noise
    return F.dropout2d(x, p, True)


# ----------------------------------------------------------------------------- image branch
def basic_block(cx, x, pre, stride):
    """resnet.py:50-66"""
    out = F.conv2d(x, cx.p(pre + ".conv1.weight"), None, stride, 1)
    out = F.relu(batch_norm(cx, out, pre + ".bn1"))
    out = F.conv2d(out, cx.p(pre + ".conv2.weight"), None, 1, 1)
    out = batch_norm(cx, out, pre + ".bn2")
    if (pre + ".downsample.0.weight") in cx.sd:
        res = F.conv2d(x, cx.p(pre + ".downsample.0.weight"), None, stride, 0)
        res = batch_norm(cx, res, pre + ".downsample.1")
    else:
        res = x
    return F.relu(out + res)


def resnet18_feats(cx, x, pre):
    """resnet.py:182-202 with layers from :153-180 — layer2 stride 2, layer3/4 stride 1 and dilation 1."""
    x = F.conv2d(x, cx.p(pre + ".conv1.weight"), None, 2, 3)
    x = F.relu(batch_norm(cx, x, pre + ".bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, stride in ((1, 1), (2, 2), (3, 1), (4, 1)):
        x = basic_block(cx, x, f"{pre}.layer{li}.0", stride)
        x = basic_block(cx, x, f"{pre}.layer{li}.1", 1)
    return x


def psp_module(cx, f, pre):
    """modules.py:27-34"""
    h, w = f.shape[2], f.shape[3]
    priors = []
    for i, s in enumerate((1, 2, 3, 6)):
        q = F.adaptive_avg_pool2d(f, (s, s))
        q = F.conv2d(q, cx.p(f"{pre}.stages.{i}.1.weight"))
        priors.append(F.interpolate(q, size=(h, w), mode="bilinear", align_corners=False))
    priors.append(f)
    return F.relu(F.conv2d(torch.cat(priors, 1), cx.p(pre + ".bottleneck.weight"), cx.p(pre + ".bottleneck.bias")))


def psp_upsample(cx, x, pre):
    """modules.py:37-48"""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = F.conv2d(x, cx.p(pre + ".conv.1.weight"), cx.p(pre + ".conv.1.bias"), 1, 1)
    x = batch_norm(cx, x, pre + ".conv.2")
    return F.prelu(x, cx.p(pre + ".conv.3.weight"))


def modified_pspnet(cx, rgb, pre):
    """modules.py:69-81 (ModifiedResnet modules.py:234-241 adds the `.model` prefix)"""
    f = resnet18_feats(cx, rgb, pre + ".feats")
    p = psp_module(cx, f, pre + ".psp")
    p = dropout2d(cx, p, 0.3)
    p = psp_upsample(cx, p, pre + ".up_1")
    p = dropout2d(cx, p, 0.15)
    p = psp_upsample(cx, p, pre + ".up_2")
    p = dropout2d(cx, p, 0.15)
    p = psp_upsample(cx, p, pre + ".up_3")
    p = F.conv2d(p, cx.p(pre + ".final.0.weight"), cx.p(pre + ".final.0.bias"))
    p = batch_norm(cx, p, pre + ".final.1")
    return F.prelu(p, cx.p(pre + ".final.2.weight"))


# ----------------------------------------------------------------------------- PointNet++ (autograd glue)
class _Group(torch.autograd.Function):
    """pointnet2_utils.py:209-257"""

    @staticmethod
    def forward(ctx, feats, idx, ops):
        ctx.ops, ctx.n = ops, feats.shape[2]
        ctx.save_for_backward(idx)
        return ops.group_points(feats.contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return ctx.ops.group_points_grad(g.contiguous(), idx, ctx.n), None, None


class _Interp(torch.autograd.Function):
    """pointnet2_utils.py:151-203"""

    @staticmethod
    def forward(ctx, feats, idx, w, ops):
        ctx.ops, ctx.m = ops, feats.shape[2]
        ctx.save_for_backward(idx, w)
        return ops.three_interpolate(feats.contiguous(), idx, w)

    @staticmethod
    def backward(ctx, g):
        idx, w = ctx.saved_tensors
        return ctx.ops.three_interpolate_grad(g.contiguous(), idx, w, ctx.m), None, None, None


def shared_mlp(cx, x, pre, nlayers):
    """pytorch_utils.py:25-50: [conv1x1(no bias) -> BN2d -> ReLU] * nlayers"""
    for i in range(nlayers):
        x = F.conv2d(x, cx.p(f"{pre}.layer{i}.conv.weight"))
        x = F.relu(batch_norm(cx, x, f"{pre}.layer{i}.normlayer.bn"))
    return x


def sa_module_msg(cx, xyz, feats, pre, npoint, radii, nsamples):
    """pointnet2_modules.py:29-73 + QueryAndGroup pointnet2_utils.py:317-377"""
    ops = cx.ops
    xyz = xyz.contiguous()
    with torch.no_grad():
        fidx = ops.furthest_point_sampling(xyz, npoint)
        xyz_t = xyz.transpose(1, 2).contiguous()
        new_xyz = ops.gather_points(xyz_t, fidx).transpose(1, 2).contiguous()
    outs = []
    for s, (r, ns) in enumerate(zip(radii, nsamples)):
        with torch.no_grad():
            idx = ops.ball_query(new_xyz, xyz, r, ns)
            g_xyz = ops.group_points(xyz_t, idx)
            g_xyz = g_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if feats is not None:
            g = torch.cat([g_xyz, _Group.apply(feats, idx, ops)], 1)
        else:
            g = g_xyz
        g = shared_mlp(cx, g, f"{pre}.mlps.{s}", 3)
        outs.append(F.max_pool2d(g, kernel_size=[1, g.shape[3]]).squeeze(-1))
    return new_xyz, torch.cat(outs, 1)


def fp_module(cx, unknown, known, unknown_feats, known_feats, pre):
    """pointnet2_modules.py:164-209"""
    ops = cx.ops
    with torch.no_grad():
        d2, idx = ops.three_nn(unknown.contiguous(), known.contiguous())
        dist = torch.sqrt(d2)
        recip = 1.0 / (dist + 1e-8)
        w = recip / torch.sum(recip, dim=2, keepdim=True)
    x = _Interp.apply(known_feats, idx, w.contiguous(), ops)
    if unknown_feats is not None:
        x = torch.cat([x, unknown_feats], 1)
    return shared_mlp(cx, x.unsqueeze(-1), pre + ".mlp", 2).squeeze(-1)


SA_NPOINT = (512, 256, 128, 64)  # modules.py:251,264,277,290
SA_NSAMPLE = (16, 32)


def pointnet2_msg(cx, pts, pre, radii_list):
    """modules.py:311-327"""
    l_xyz, l_f = [pts.contiguous()], [None]
    for i in range(4):
        nx, nf = sa_module_msg(cx, l_xyz[i], l_f[i], f"{pre}.SA_modules.{i}", SA_NPOINT[i], radii_list[i], SA_NSAMPLE)
        l_xyz.append(nx)
        l_f.append(nf)
    for i in range(-1, -5, -1):
        l_f[i - 1] = fp_module(cx, l_xyz[i - 1], l_xyz[i], l_f[i - 1], l_f[i], f"{pre}.FP_modules.{4 + i}")
    return l_f[0]


CAM_RADII = [[0.01, 0.02], [0.02, 0.04], [0.04, 0.08], [0.08, 0.16]]  # ist_net.py:16
WORLD_RADII = [[0.05, 0.10], [0.10, 0.20], [0.20, 0.30], [0.30, 0.40]]  # ist_net.py:189


# ----------------------------------------------------------------------------- per-point MLP blocks
def conv1d_seq(cx, x, pre, idxs, last_relu=True):
    for j, i in enumerate(idxs):
        x = F.conv1d(x, cx.p(f"{pre}.{i}.weight"), cx.p(f"{pre}.{i}.bias"))
        if last_relu or j + 1 < len(idxs):
            x = F.relu(x)
    return x


def linear_head(cx, x, pre):
    x = F.relu(F.linear(x, cx.p(pre + ".0.weight"), cx.p(pre + ".0.bias")))
    x = F.relu(F.linear(x, cx.p(pre + ".2.weight"), cx.p(pre + ".2.bias")))
    return F.linear(x, cx.p(pre + ".4.weight"), cx.p(pre + ".4.bias"))


def ortho6d_to_mat(x_raw, y_raw):
    """utils/rotation_utils.py:4-28"""

    def nrm(v):
        mag = torch.sqrt(v.pow(2).sum(dim=1, keepdim=True))
        # the reference builds this constant on the host each call (rotation_utils.py:6: FloatTensor([1e-8]).cuda(), a synchronous
        # copy that cannot be captured in a CUDA graph); new_full is the same value created on the tensor's device
        return v / torch.max(mag, mag.new_full((1,), 1e-8))

    def cross(u, v):
        return torch.stack(
            (u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1], u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2], u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]), 1
        )

    y = nrm(y_raw)
    z = nrm(cross(x_raw, y))
    x = cross(y, z)
    return torch.stack((x, y, z), 2)


def _estimator_tail(cx, feat, pre):
    """shared tail of ist_net.py:250-264 and :318-332"""
    feat = conv1d_seq(cx, feat, pre + ".pose_mlp1", (0, 2))
    glob = torch.mean(feat, 2, keepdim=True)
    feat = torch.cat([feat, glob.expand_as(feat)], 1)
    feat = conv1d_seq(cx, feat, pre + ".pose_mlp2", (0, 2))
    feat = F.adaptive_avg_pool1d(feat, 1).squeeze(2)
    r6 = linear_head(cx, feat, pre + ".rotation_estimator")
    r = ortho6d_to_mat(r6[:, :3].contiguous(), r6[:, 3:].contiguous()).view(-1, 3, 3)
    return r, linear_head(cx, feat, pre + ".translation_estimator"), linear_head(cx, feat, pre + ".size_estimator")


def light_estimator(cx, pts, rgb_local, pts_local, pre):
    """ist_net.py:250-264"""
    e = conv1d_seq(cx, pts.transpose(1, 2), pre + ".pts_mlp", (0, 2))
    return _estimator_tail(cx, torch.cat([rgb_local, e, pts_local], 1), pre)


def heavy_estimator(cx, pts, pts_w, rgb_local, pts_local, pts_w_local, pre):
    """ist_net.py:318-332 (duplicate at posenet_gt.py:122-136)"""
    e1 = conv1d_seq(cx, pts.transpose(1, 2), pre + ".pts_mlp1", (0, 2))
    e2 = conv1d_seq(cx, pts_w.transpose(1, 2), pre + ".pts_mlp2", (0, 2))
    return _estimator_tail(cx, torch.cat([rgb_local, e1, pts_local, e2, pts_w_local], 1), pre)


def feature_deformer(cx, pts, rgb_local, pts_local, index, pre):
    """ist_net.py:162-183"""
    n = pts_local.shape[2]
    e = conv1d_seq(cx, pts.transpose(1, 2), pre + ".pts_mlp1", (0, 2))
    x = conv1d_seq(cx, torch.cat([e, pts_local, rgb_local], 1), pre + ".deform_mlp1", (0, 2))
    glob = torch.mean(x, 2, keepdim=True)
    x = conv1d_seq(cx, torch.cat([x, glob.expand_as(x)], 1), pre + ".deform_mlp2", (0, 2, 4))
    q = conv1d_seq(cx, x, pre + ".pred_nocs", (0, 2, 4), last_relu=False)
    q = q.view(-1, 3, n).contiguous()
    q = torch.index_select(q, 0, index).permute(0, 2, 1).contiguous()
    return x, q


def gather_pixels(rgb_map, choose):
    """ist_net.py:42-45"""
    b, d = rgb_map.shape[:2]
    flat = rgb_map.view(b, d, -1)
    return torch.gather(flat, 2, choose.unsqueeze(1).repeat(1, d, 1)).contiguous()


# ----------------------------------------------------------------------------- top-level models
def ist_net_forward(sd, inputs, training, nclass=6, freeze_world_enhancer=False, **kw):
    """IST_Net.forward, ist_net.py:22-76"""
    cx = Ctx(sd, training, **kw)
    pts = inputs["pts"]
    cls = inputs["category_label"].reshape(-1)
    c = torch.mean(pts, 1, keepdim=True)
    pts = pts - c
    b = pts.shape[0]
    index = cls + torch.arange(b, dtype=torch.long, device=pts.device) * nclass
    rgb_local = gather_pixels(modified_pspnet(cx, inputs["rgb"], "rgb_cam_extractor.model"), inputs["choose"])
    ep = {}
    pts_local = pointnet2_msg(cx, pts, "pts_cam_extractor", CAM_RADII)
    if training:
        r_c, t_c, s_c = light_estimator(cx, pts, rgb_local, pts_local, "cam_enhancer")
    pts_w_local, pts_w = feature_deformer(cx, pts, rgb_local, pts_local, index, "implicit_transform.feature_refine")
    r, t, s = heavy_estimator(cx, pts, pts_w, rgb_local, pts_local, pts_w_local, "main_estimator")
    ep["pred_qo"] = pts_w
    ep["pred_rotation"], ep["pred_translation"], ep["pred_size"] = r, t + c.squeeze(1), s
    if training:
        qo = inputs["qo"]
        pts_w_local_gt = pointnet2_msg(cx, qo, "world_enhancer.extractor", WORLD_RADII)  # ist_net.py:193-200
        ep["pts_w_local"], ep["pts_w_local_gt"] = pts_w_local, pts_w_local_gt
        ep["pred_rotation_aux_cam"], ep["pred_translation_aux_cam"], ep["pred_size_aux_cam"] = r_c, t_c + c.squeeze(1), s_c
        if not freeze_world_enhancer:
            r_w, t_w, s_w = heavy_estimator(
                cx, pts, qo, rgb_local.detach(), pts_local.detach(), pts_w_local_gt, "world_enhancer.pose_estimator"
            )
            ep["pred_rotation_aux_world"], ep["pred_translation_aux_world"], ep["pred_size_aux_world"] = r_w, t_w + c.squeeze(1), s_w
    return ep


def posenet_gt_forward(sd, inputs, training, **kw):
    """PoseNetGT.forward, posenet_gt.py:22-51"""
    cx = Ctx(sd, training, **kw)
    pts = inputs["pts"]
    c = torch.mean(pts, 1, keepdim=True)
    pts = pts - c
    rgb_local = gather_pixels(modified_pspnet(cx, inputs["rgb"], "rgb_extractor.model"), inputs["choose"])
    pts_local = pointnet2_msg(cx, pts, "pts_extractor", CAM_RADII)
    gt_local = pointnet2_msg(cx, inputs["qo"], "pts_gt_extractor", WORLD_RADII)
    r, t, s = heavy_estimator(cx, pts, inputs["qo"], rgb_local.detach(), pts_local.detach(), gt_local, "pose_estimator_aux")
    return {"pts_local_w_gt": gt_local, "pred_rotation": r, "pred_translation": t + c.squeeze(1), "pred_size": s}


# ----------------------------------------------------------------------------- losses
def smooth_l1_dis(p1, p2, threshold=0.1):
    """losses.py:3-22"""
    diff = torch.abs(p1 - p2)
    dis = torch.where(diff > threshold, diff - threshold / 2.0, torch.pow(diff, 2) / (2.0 * threshold))
    return torch.mean(torch.sum(dis, dim=2 if p1.dim() == 3 else 1))


def pose_dis(r1, t1, s1, r2, t2, s2):
    """losses.py:37-49"""
    return torch.mean(torch.norm(r1 - r2, dim=1)) + torch.mean(torch.norm(t1 - t2, dim=1)) + torch.mean(torch.norm(s1 - s2, dim=1))


def ist_net_loss(ep, labels, gamma1=1.0, gamma2=10.0, freeze_world_enhancer=False):
    """SupervisedLoss, ist_net.py:78-111 (gammas: config/ist_net_default.yaml:28-30)"""
    R, T, S = labels["rotation_label"], labels["translation_label"], labels["size_label"]
    loss = pose_dis(ep["pred_rotation"], ep["pred_translation"], ep["pred_size"], R, T, S)
    loss = loss + pose_dis(ep["pred_rotation_aux_cam"], ep["pred_translation_aux_cam"], ep["pred_size_aux_cam"], R, T, S)
    loss = loss + gamma1 * smooth_l1_dis(ep["pred_qo"], labels["qo"]) + gamma2 * F.mse_loss(ep["pts_w_local"], ep["pts_w_local_gt"])
    if not freeze_world_enhancer:
        loss = loss + pose_dis(ep["pred_rotation_aux_world"], ep["pred_translation_aux_world"], ep["pred_size_aux_world"], R, T, S)
    return loss


def posenet_gt_loss(ep, labels):
    """posenet_gt.py:53-67"""
    return pose_dis(ep["pred_rotation"], ep["pred_translation"], ep["pred_size"], labels["rotation_label"], labels["translation_label"], labels["size_label"])
