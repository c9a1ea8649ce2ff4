"""PointNet++ MSG encoder of IST-Net on the B200 kernels.

Module / parameter names mirror the reference so `state_dict`s interchange
(model/modules.py:244-327 `PointNet2MSG`, model/pointnet2/pointnet2_modules.py:21-209,
model/pointnet2/pytorch_utils.py:25-206 `SharedMLP`): e.g. `SA_modules.0.mlps.0.layer0.conv.weight`,
`SA_modules.0.mlps.0.layer0.normlayer.bn.running_var`, `FP_modules.3.mlp.layer1.conv.weight`.
BN layers are real `nn.BatchNorm2d` instances so the reference BNMomentumScheduler (utils/scheduler.py:277-303)
finds them; momentum / eps / training are read at call time.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ext
from . import functional as PF
from . import rows_engine as RE
from . import sa_fused as SF

SA_FORK = os.environ.get("ISTNET_SA_FORK", "1") != "0"  # the two scales of an unfused SA level on two streams


class _NormLayer(nn.Sequential):
    def __init__(self, c):
        super().__init__()
        self.add_module("bn", nn.BatchNorm2d(c))
        nn.init.constant_(self.bn.weight, 1.0)
        nn.init.constant_(self.bn.bias, 0.0)


class _ConvBnRelu(nn.Sequential):
    """pytorch_utils.Conv2d with bn=True: 1x1 conv without bias, BN2d, ReLU (pytorch_utils.py:80-134,173-206)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.add_module("conv", nn.Conv2d(cin, cout, kernel_size=(1, 1), bias=False))
        nn.init.kaiming_normal_(self.conv.weight)
        self.add_module("normlayer", _NormLayer(cout))
        self.add_module("activation", nn.ReLU(inplace=True))


class SharedMLP(nn.Sequential):
    def __init__(self, spec, bn=True):
        super().__init__()
        assert bn, "the IST-Net hot path only uses bn=True"
        for i in range(len(spec) - 1):
            self.add_module(f"layer{i}", _ConvBnRelu(spec[i], spec[i + 1]))


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:294-377 (use_xyz=True, no normalize_xyz / sample_uniformly on this path).  Holds the ball
    parameters; the grouping itself is fused into the first shared-MLP operand (rows_engine.sa_scale)."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        """Reference-layout result (B, 3+C, npoint, nsample) — API compatibility, not used by the fused path."""
        idx = PF.ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_t = xyz.transpose(1, 2).contiguous()
        with torch.no_grad():
            grouped_xyz = ext.group_points(xyz_t, idx)
            grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            return grouped_xyz
        grouped = PF.group_points(features, idx)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped


class PointnetSAModuleMSG(nn.Module):
    """pointnet2_modules.py:21-114"""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps) and use_xyz
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for r, ns, spec in zip(radii, nsamples, mlps):
            self.groupers.append(QueryAndGroup(r, ns, use_xyz=use_xyz))
            spec = list(spec)
            spec[0] += 3
            self.mlps.append(SharedMLP(spec, bn=bn))

    def forward_rows(self, xyz, feats_rows=None, new_xyz=None):
        """xyz (B,N,3), feats_rows (B,N,C) channels-last -> new_xyz (B,npoint,3), new feats (B,npoint,sum mlp[-1])."""
        xyz = xyz.contiguous()
        if new_xyz is None:
            with torch.no_grad():
                _, cent = ext.fps_chain(xyz, (self.npoint,))
            new_xyz = cent[0]
        if SF.supported(self, xyz, new_xyz, feats_rows):
            bn0 = self.mlps[0][0].normlayer.bn
            needs_grad = torch.is_grad_enabled() and ((feats_rows is not None and feats_rows.requires_grad) or any(p.requires_grad for p in self.parameters()))
            if RE.K.bn_uses_batch_stats(bn0, self.training) or not needs_grad:
                # both scales, ball query + grouping + SharedMLP + max in one launch per pass (csrc/sa_fused.cu)
                return new_xyz, SF.sa_level(self, xyz, new_xyz, feats_rows)
        def scale(s):
            grouper, mlp = self.groupers[s], self.mlps[s]
            idx = PF.ball_query(grouper.radius, grouper.nsample, xyz, new_xyz)
            return RE.sa_scale(RE.units_from_shared_mlp(mlp), self.training, xyz, new_xyz, idx, feats_rows)

        # the scales of a level are independent chains of small launches (levels 3-4: 64..256 centroids per instance): the last scale
        # runs on a side stream of the current one, forward and (autograd replays a node on its forward stream) backward
        side = RE.K.side_stream_for(xyz.device, 1) if (SA_FORK and xyz.is_cuda and len(self.groupers) > 1) else None
        if side is None:
            return new_xyz, torch.cat([scale(s) for s in range(len(self.groupers))], dim=2)
        main = torch.cuda.current_stream(xyz.device)
        side.wait_stream(main)
        last = len(self.groupers) - 1
        with torch.cuda.stream(side):
            o_last = scale(last)
        outs = [scale(s) for s in range(last)] + [o_last]
        main.wait_stream(side)
        o_last.record_stream(main)
        return new_xyz, torch.cat(outs, dim=2)

    def forward(self, xyz, features=None, new_xyz=None):
        """Reference layout (pointnet2_modules.py:29-73): features (B,C,N) -> (B,sum mlp[-1],npoint)."""
        fr = features.transpose(1, 2).contiguous() if features is not None else None
        new_xyz, out = self.forward_rows(xyz, fr, new_xyz)
        return new_xyz, out.transpose(1, 2).contiguous()


class PointnetFPModule(nn.Module):
    """pointnet2_modules.py:153-209"""

    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = SharedMLP(mlp, bn=bn)

    def forward_rows(self, unknown, known, unknown_rows, known_rows):
        """unknown (B,n,3), known (B,m,3), unknown_rows (B,n,C1) or None, known_rows (B,m,C2) -> (B,n,mlp[-1])"""
        out = RE.fp_level(RE.units_from_shared_mlp(self.mlp), self.training, unknown, known, known_rows, unknown_rows)
        if out is not None:
            return out
        idx, weight = PF.three_nn_weights(unknown, known)
        x = RE.interp_rows(known_rows, idx, weight)
        B, n, _ = x.shape
        srcs = [x.reshape(B * n, -1)]
        if unknown_rows is not None:  # torch.cat([interpolated, skip], dim=1) of pointnet2_modules.py:196-199 happens in the operand split
            srcs.append(unknown_rows.reshape(B * n, -1))
        return RE.run_chain(RE.units_from_shared_mlp(self.mlp), srcs, self.training).view(B, n, -1)

    def forward(self, unknown, known, unknow_feats, known_feats):
        ur = unknow_feats.transpose(1, 2).contiguous() if unknow_feats is not None else None
        out = self.forward_rows(unknown, known, ur, known_feats.transpose(1, 2).contiguous())
        return out.transpose(1, 2).contiguous()


class PointNet2MSG(nn.Module):
    """modules.py:244-327: 4 SA-MSG levels (512/256/128/64 centroids, 16|32 neighbours) + 4 FP levels."""

    NPOINT = (512, 256, 128, 64)

    def __init__(self, radii_list, use_xyz=True):
        super().__init__()
        self.SA_modules = nn.ModuleList()
        c_in, widths = 0, ((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128, 256))
        outs = []
        for lvl in range(4):
            w = widths[lvl]
            self.SA_modules.append(
                PointnetSAModuleMSG(self.NPOINT[lvl], radii_list[lvl], [16, 32], [[c_in, *w], [c_in, *w]], use_xyz=use_xyz, bn=True)
            )
            c_in = 2 * w[-1]
            outs.append(c_in)
        self.FP_modules = nn.ModuleList()
        self.FP_modules.append(PointnetFPModule(mlp=[256, 128, 128]))
        self.FP_modules.append(PointnetFPModule(mlp=[256 + outs[0], 256, 256]))
        self.FP_modules.append(PointnetFPModule(mlp=[512 + outs[1], 256, 256]))
        self.FP_modules.append(PointnetFPModule(mlp=[outs[3] + outs[2], 512, 512]))

    def forward_rows(self, pointcloud):
        """(B,N,3[+C]) -> per-point features as rows (B,N,128)."""
        xyz = pointcloud[..., 0:3].contiguous()
        feats = pointcloud[..., 3:].contiguous() if pointcloud.size(-1) > 3 else None
        # all four FPS levels + centroid gathers in ONE launch (the reference runs 4 FPS + 4 gather kernels)
        with torch.no_grad():
            _, centroids = ext.fps_chain(xyz, self.NPOINT)
        l_xyz, l_feats = [xyz], [feats]
        for i, sa in enumerate(self.SA_modules):
            nx, nf = sa.forward_rows(l_xyz[i], l_feats[i], new_xyz=centroids[i])
            l_xyz.append(nx)
            l_feats.append(nf)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_feats[i - 1] = self.FP_modules[i].forward_rows(l_xyz[i - 1], l_xyz[i], l_feats[i - 1], l_feats[i])
        return l_feats[0]

    def forward(self, pointcloud):
        """Reference layout (modules.py:311-327): (B,128,N)."""
        return self.forward_rows(pointcloud).transpose(1, 2).contiguous()
