"""Network mirrors of the reference's network/upsampler.py: Net (:9-189) and Level (:192-374).

Same constructor arguments, module/parameter names (state_dict keys `levels.level_{l}....`, SURVEY.md
appendix A) and the same results, on libpu3_b200 kernels.  What differs from the reference:
  * eval-mode Net.forward accepts a BATCH of input patches.  The reference asserts batch_size == 1 past
    level 1 (upsampler.py:61) and main.py loops over patches (main.py:237-244); here the B patches run
    together and every one of them gets exactly the result of its own B=1 call.
  * Level.forward writes every layer's output straight into its slot of one (B,264,N) feature buffer
    instead of torch.cat-ing (upsampler.py:293,299,305,311), and the neighbourhood tensors of
    DenseEdgeConv never exist (fused kernel).
AdaptiveLevel (:377-512, never instantiated) and the phase == "vis" debug branches are not provided.
"""
from collections import OrderedDict
from math import log, sqrt

import torch

from . import fused, layers, level_train, operations


class Level(torch.nn.Module):
    """3PU per-level network (upsampler.py:192-374)."""

    use_engine = True    # eval fast path through the C++ level engine (False: one Python call per kernel, for profiling)

    def __init__(self, dense_n=3, growth_rate=12, knn=16, fm_knn=5, step_ratio=2):
        super(Level, self).__init__()
        self.dense_n = dense_n
        self.fm_knn = fm_knn
        self.step_ratio = step_ratio
        self.knn = knn
        self.growth_rate = growth_rate
        if step_ratio < 4:
            self.code = self.gen_1d_grid(step_ratio).unsqueeze(0).detach()      # 1 x 1 x r
        else:
            expansion_ratio = round(sqrt(step_ratio)) ** 2
            self.code = self.gen_grid(expansion_ratio).unsqueeze(0).detach()    # 1 x 2 x r
        self._code_dev = {}

        comp = 24 + growth_rate * dense_n                                       # channels a dense block adds
        self.layer0 = layers.Conv2d(3, 24, [1, 1], activation=None)
        self.layer1 = layers.DenseEdgeConv(24, growth_rate=growth_rate, n=dense_n, k=knn)
        in_channels = 24 + comp
        self.layer2_prep = layers.Conv1d(in_channels, 24, 1, activation="relu")
        self.layer2 = layers.DenseEdgeConv(24, growth_rate=growth_rate, n=dense_n, k=knn)
        in_channels += comp
        self.layer3_prep = layers.Conv1d(in_channels, 24, 1, activation="relu")
        self.layer3 = layers.DenseEdgeConv(24, growth_rate=growth_rate, n=dense_n, k=knn)
        in_channels += comp
        self.layer4_prep = layers.Conv1d(in_channels, 24, 1, activation="relu")
        self.layer4 = layers.DenseEdgeConv(24, growth_rate=growth_rate, n=dense_n, k=knn)
        in_channels += comp
        self.feat_channels = in_channels
        self.up_layer = torch.nn.Sequential(OrderedDict([
            ("up_layer1", layers.Conv2d(in_channels + self.code.size(1), 128, 1, activation="relu")),
            ("up_layer2", layers.Conv2d(128, 128, 1, activation="relu")), ]))
        self.fc_layer1 = layers.Conv2d(128, 64, 1, activation="relu")
        self.fc_layer2 = layers.Conv2d(64, 3, 1, activation=None)

    # ---- helpers kept from the reference ---------------------------------------------------------
    def exponential_distance(self, points, knnIdx_points):
        """upsampler.py:232-250: squared distance to the k neighbours and exp(-d / (h/2)), h = mean_N(min_K d)."""
        if points.dim() == 3:
            points = points.unsqueeze(dim=-1)
        distance = torch.sum((points - knnIdx_points) ** 2, dim=1, keepdim=True).detach()
        h = torch.mean(torch.min(distance, dim=-1, keepdim=True)[0], dim=-2, keepdim=True)
        weight = torch.exp(-distance / (h / 2)).detach()
        return distance, weight

    def gen_grid(self, grid_size):
        """output [2, grid_size x grid_size] (upsampler.py:252-262)"""
        x = torch.linspace(-0.2, 0.2, grid_size, dtype=torch.float32)
        x, y = torch.meshgrid(x, x, indexing="ij")
        return torch.stack([x, y], dim=0).view([2, grid_size * grid_size])

    def gen_1d_grid(self, num_grid_point):
        """output [1, num_grid_point] (upsampler.py:264-270)"""
        return torch.linspace(-0.2, 0.2, num_grid_point).view(1, num_grid_point)

    def _code_on(self, device):
        c = self._code_dev.get(device)
        if c is None:
            c = self.code.to(device=device, dtype=torch.float32).contiguous()
            self._code_dev[device] = c
        return c

    def _code_row(self, device, T, N):
        """(T,1,N*r) tensor holding code[p % r]: the code channel of the expanded tensor (upsampler.py:354-361), for its weight gradient"""
        key = (str(device), T, N)
        cache = self.__dict__.setdefault("_code_rows", {})
        if key not in cache:
            cache.clear()
            cache[key] = self._code_on(device).view(1, 1, -1).repeat(T, 1, N).contiguous()
        return cache[key]

    def _engine_params(self):
        """the 40 parameters in the order of pu3_level_weights"""
        params = [self.layer0.conv.weight, self.layer0.conv.bias]
        for blk in (self.layer1, self.layer2, self.layer3, self.layer4):
            for m in blk.mlps:
                params += [m.weight, m.bias]
        for prep in (self.layer2_prep, self.layer3_prep, self.layer4_prep):
            params += [prep.conv.weight, prep.conv.bias]
        up1, up2 = self.up_layer.up_layer1.conv, self.up_layer.up_layer2.conv
        params += [up1.weight, up1.bias, up2.weight, up2.bias, self.fc_layer1.conv.weight, self.fc_layer1.conv.bias,
                   self.fc_layer2.conv.weight, self.fc_layer2.conv.bias]
        return params

    # train mode as ONE autograd node on native forward + backward kernels (level_train.py); False: the operator composition
    # below (every layer its own autograd function, skip connection through torch operators) -- kept for tests
    native_train = True

    def _native_train_ok(self, xyz_normalized, previous_level4):
        n = xyz_normalized.shape[2]
        ok = (xyz_normalized.is_cuda and xyz_normalized.dtype == torch.float32 and self.dense_n == 3 and self.growth_rate == 12
              and self.code.size(1) == 1 and self._engine_ok() and self.knn <= 32 and n % 4 == 0 and n <= 1500
              and self.code.size(2) <= 8 and n >= self.knn + 1)
        if ok and previous_level4 is not None and self.fm_knn > 0:
            pf = previous_level4[1]
            ok = pf.dim() == 3 and pf.shape[1] == self.feat_channels and xyz_normalized.shape[0] % pf.shape[0] == 0
        return ok

    # ---- forward ---------------------------------------------------------------------------------------
    def _fast_path_ok(self, xyz_normalized):
        # gradients wanted (training mode, or an input that requires grad): differentiable composition instead
        wants_grad = torch.is_grad_enabled() and (self.training or xyz_normalized.requires_grad)
        return (xyz_normalized.is_cuda and xyz_normalized.dtype == torch.float32 and self.dense_n == 3
                and self.growth_rate == 12 and self.code.size(1) == 1 and not wants_grad)

    # ---- level engine: one C call for the whole forward (csrc/level.cu) --------------------------------------
    def _engine_weights(self, device):
        """ctypes pu3_level_weights for the current parameter storage (rebuilt when a parameter moved or changed)."""
        params = [self.layer0.conv.weight, self.layer0.conv.bias]
        for blk in (self.layer1, self.layer2, self.layer3, self.layer4):
            for m in blk.mlps:
                params += [m.weight, m.bias]
        for prep in (self.layer2_prep, self.layer3_prep, self.layer4_prep):
            params += [prep.conv.weight, prep.conv.bias]
        up1, up2 = self.up_layer.up_layer1.conv, self.up_layer.up_layer2.conv
        params += [up1.weight, up1.bias, up2.weight, up2.bias, self.fc_layer1.conv.weight, self.fc_layer1.conv.bias,
                   self.fc_layer2.conv.weight, self.fc_layer2.conv.bias]
        key = tuple((p.data_ptr(), p._version) for p in params) + (str(device), fused._lib.weight_generation)
        cached = self.__dict__.get("_engine_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        if not all(p.is_contiguous() and p.device == device and p.dtype == torch.float32 for p in params):
            raise RuntimeError("Level parameters must be contiguous float32 tensors on the input's device")
        W = fused._lib.LevelWeights()
        it = iter(params)
        W.layer0_w, W.layer0_b = next(it).data_ptr(), next(it).data_ptr()
        for bi in range(4):
            for li in range(3):
                W.ec_w[bi][li], W.ec_b[bi][li] = next(it).data_ptr(), next(it).data_ptr()
        for pi in range(3):
            W.prep_w[pi], W.prep_b[pi] = next(it).data_ptr(), next(it).data_ptr()
        C = self.feat_channels
        up1_feat = up1.weight.detach().reshape(128, C + 1)[:, :C].contiguous()     # without the code column
        code = self._code_on(device)
        W.up1_w, W.up1_w_feat, W.up1_b = up1.weight.data_ptr(), up1_feat.data_ptr(), up1.bias.data_ptr()
        W.up2_w, W.up2_b = up2.weight.data_ptr(), up2.bias.data_ptr()
        W.fc1_w, W.fc1_b = self.fc_layer1.conv.weight.data_ptr(), self.fc_layer1.conv.bias.data_ptr()
        W.fc2_w, W.fc2_b = self.fc_layer2.conv.weight.data_ptr(), self.fc_layer2.conv.bias.data_ptr()
        W.code = code.data_ptr()
        W.r, W.knn, W.fm_knn, W.reserved = int(self.code.size(2)), int(self.knn), int(self.fm_knn), 0
        self.__dict__["_engine_cache"] = (key, W, up1_feat, code)       # keep the derived tensors alive
        return W

    def _engine_ok(self):
        return (self.feat_channels == 264 and all(b.k == self.knn for b in (self.layer1, self.layer2, self.layer3, self.layer4))
                and self.up_layer.up_layer1.conv.weight.shape[1] == 265)

    def _forward_engine(self, xyz, xyz_normalized, previous_level4, group, ragged, feat_pm_out=None):
        import ctypes
        T, _, N = xyz_normalized.shape
        dev = xyz_normalized.device
        W = self._engine_weights(dev)
        r = W.r
        xn = xyz_normalized.contiguous()
        has_prev = previous_level4 is not None and self.fm_knn > 0
        if has_prev:
            prev_xyz, prev_feat_pm = previous_level4
            prev_xyz, prev_feat_pm = prev_xyz.contiguous(), prev_feat_pm.contiguous()
            clouds, No = prev_feat_pm.shape[0], prev_feat_pm.shape[1]
            xyz_c = xyz.contiguous()
        else:
            prev_xyz = prev_feat_pm = xyz_c = None
            clouds, No = 0, 0
        owner = ragged.owner if ragged is not None else None
        groups = ragged.groups if ragged is not None else 0
        prev_n = ragged.n_arr if (ragged is not None and has_prev) else None
        L = fused._lib.lib()
        ws_bytes = L.pu3_level_workspace(T, N, r, W.knn, W.fm_knn, max(clouds, 1), max(No, 1), int(has_prev))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        feat = torch.empty(T, 264, N, dtype=torch.float32, device=dev)
        out = torch.empty(T, 3, N * r, dtype=torch.float32, device=dev)
        extra = (5 if has_prev else 0) + (1 if owner is not None else 0)
        if feat_pm_out is not None:
            # the features once more point-major (T,N,264) for the next level's skip connection: written by this level's skip
            # kernel while the rows are in registers (level 1: by a transposing pass inside the engine)
            assert feat_pm_out.is_contiguous() and feat_pm_out.numel() == T * N * 264
            fused._lib.launch("pu3_level_forward_pm_f32", xn, ctypes.addressof(W), T, N, fused._lib.ptr(xyz_c), xn.data_ptr(),
                              fused._lib.ptr(owner), int(groups), int(group or T), fused._lib.ptr(prev_xyz),
                              fused._lib.ptr(prev_feat_pm), int(clouds), int(No), fused._lib.ptr(prev_n), feat.data_ptr(),
                              out.data_ptr(), feat_pm_out.data_ptr(), ws.data_ptr(), ws_bytes,
                              extra_kernels=extra + (0 if has_prev else 1), tag="pu3_level_forward_f32")
            return out, feat
        fused._lib.launch("pu3_level_forward_f32", xn, ctypes.addressof(W), T, N, fused._lib.ptr(xyz_c), xn.data_ptr(),
                          fused._lib.ptr(owner), int(groups), int(group or T), fused._lib.ptr(prev_xyz),
                          fused._lib.ptr(prev_feat_pm), int(clouds), int(No), fused._lib.ptr(prev_n), feat.data_ptr(),
                          out.data_ptr(), ws.data_ptr(), ws_bytes, extra_kernels=extra)
        return out, feat

    def _features_fused(self, xyz_normalized, group, ragged=None):
        """layer0 + 4 dense blocks into one (B,264,N) buffer; channel order [y4, y3, y2, y1, x0]."""
        B, _, N = xyz_normalized.shape
        self_ragged = None
        if ragged is not None:   # every patch is its own cloud; the duplicate-penalty group is its request
            me = torch.arange(B, dtype=torch.int32, device=xyz_normalized.device)
            self_ragged = operations.Ragged(me, ragged.owner, ragged.groups)
        C = self.feat_channels
        feat = torch.empty(B, C, N, dtype=torch.float32, device=xyz_normalized.device)
        # h: the 24-channel input of the current dense block (contiguous: kNN and edge-conv read it)
        h = torch.empty(B, 24, N, dtype=torch.float32, device=feat.device)
        fused.conv_into(xyz_normalized.contiguous(), self.layer0.conv.weight, self.layer0.conv.bias, h)
        feat[:, C - 24:].copy_(h)
        lo = C - 24
        for li, (prep, block) in enumerate(((None, self.layer1), (self.layer2_prep, self.layer2),
                                            (self.layer3_prep, self.layer3), (self.layer4_prep, self.layer4))):
            if prep is not None:
                fused.conv_into(feat[:, lo:], prep.conv.weight, prep.conv.bias, h, relu=True)
            src = h
            _, idx32, _ = operations._knn_raw(block.k + 1, src, src, True, group, want_knn=False, want_dist=False,
                                              idx_dtype=torch.int32, ragged=self_ragged)
            fused.edgeconv_into(src, idx32, 1, block.k, [m.weight for m in block.mlps], [m.bias for m in block.mlps],
                                feat[:, lo - 60:lo])
            lo -= 60
        return feat

    def _skip_connection_fused(self, x, xyz, previous_level4, group, ragged=None):
        """The same computation as _skip_connection in one kernel (csrc/skip.cu), x (T,C,N) updated in place.
        previous_level4 = (prev_xyz (clouds,3,No) channel-major, prev_feat (clouds,No,C) POINT-major)."""
        previous_xyz, previous_feat_pm = previous_level4
        T, C, N = x.shape
        clouds, No, _ = previous_feat_pm.shape
        _, idx, _ = operations._knn_raw(self.fm_knn, xyz.contiguous(), previous_xyz.contiguous(), True, group,
                                        want_knn=False, want_dist=False, ragged=ragged)
        owner = ragged.owner if ragged is not None else None
        fused._lib.launch("pu3_skip_fuse_f32", x, T, N, C, self.fm_knn, 1 if ragged is not None else T // clouds, No,
                          x.data_ptr(), xyz.contiguous().data_ptr(), idx.data_ptr(), previous_xyz.data_ptr(),
                          previous_feat_pm.data_ptr(), fused._lib.ptr(owner))
        return x

    def _skip_connection(self, x, xyz, previous_level4, group, ragged=None):
        """upsampler.py:317-347: bilateral (spatial x feature) interpolation of the previous level's features."""
        previous_xyz, previous_feat = previous_level4
        B, _, N = xyz.shape
        Bp = previous_xyz.shape[0]
        if ragged is not None:
            # patch i reads the previous-level cloud of its request owner[i], which holds n_arr[owner[i]] points
            knnIdx_points, knnIdx_idx, _ = operations._knn_raw(self.fm_knn, xyz.contiguous(), previous_xyz.contiguous(),
                                                               True, None, want_dist=False, ragged=ragged)
            Cp, Mp = previous_feat.shape[1], previous_feat.shape[2]
            flat = previous_feat.permute(1, 0, 2).reshape(Cp, Bp * Mp)
            offs = (ragged.owner.long() * Mp).view(B, 1, 1)
            g = flat[:, (knnIdx_idx + offs).reshape(-1)]
            knnIdx_feats = g.view(Cp, B, N, self.fm_knn).permute(1, 0, 2, 3)
        else:
            knnIdx_points, knnIdx_idx, _ = operations.group_knn(self.fm_knn, xyz, previous_xyz, unique=True, NCHW=True,
                                                                max_group=group)
        if ragged is not None:
            pass
        elif Bp == B:
            pf = previous_feat.unsqueeze(2).expand(-1, -1, N, -1)
            knnIdx_feats = torch.gather(pf, 3, knnIdx_idx.unsqueeze(1).expand(-1, pf.size(1), -1, -1))
        else:
            # previous level shared by B/Bp consecutive patches (the reference expand()s it, :319-323)
            p_div = B // Bp
            Cp, Mp = previous_feat.shape[1], previous_feat.shape[2]
            flat = previous_feat.permute(1, 0, 2).reshape(Cp, Bp * Mp)
            offs = (torch.arange(B, device=xyz.device) // p_div * Mp).view(B, 1, 1)
            g = flat[:, (knnIdx_idx + offs).reshape(-1)]                       # Cp, B*N*K
            knnIdx_feats = g.view(Cp, B, N, self.fm_knn).permute(1, 0, 2, 3)
        _, s_average_weight = self.exponential_distance(xyz, knnIdx_points)
        _, f_average_weight = self.exponential_distance(x, knnIdx_feats)
        average_weight = s_average_weight * f_average_weight
        average_weight = average_weight / torch.sum(average_weight + 1e-5, dim=-1, keepdim=True)
        knnIdx_feats = torch.sum(average_weight * knnIdx_feats, dim=-1)
        return 0.2 * knnIdx_feats + x

    def _head_fused(self, x, xyz_normalized):
        """upsampler.py:349-372 without materialising the (B,265,N*r) replicated tensor."""
        B, C, N = x.shape
        r = self.code.size(2)
        dev = x.device
        w1 = self.up_layer.up_layer1.conv.weight.reshape(128, C + 1)
        pre = torch.empty(B, 128, N, dtype=torch.float32, device=dev)
        fused.conv_into(x.contiguous(), w1[:, :C].contiguous(), self.up_layer.up_layer1.conv.bias, pre)
        h1 = torch.empty(B, 128, N * r, dtype=torch.float32, device=dev)
        code = self._code_on(dev)
        w1c = w1.contiguous()
        fused._lib.launch("pu3_expand_code_f32", x, B, 128, N, r, pre.data_ptr(), w1c.data_ptr(), C + 1, C,
                                                                 code.data_ptr(), h1.data_ptr())
        h2 = torch.empty_like(h1)
        fused.conv_into(h1, self.up_layer.up_layer2.conv.weight, self.up_layer.up_layer2.conv.bias, h2, relu=True)
        h3 = torch.empty(B, 64, N * r, dtype=torch.float32, device=dev)
        fused.conv_into(h2, self.fc_layer1.conv.weight, self.fc_layer1.conv.bias, h3, relu=True)
        out = torch.empty(B, 3, N * r, dtype=torch.float32, device=dev)
        fused.conv_into(h3, self.fc_layer2.conv.weight, self.fc_layer2.conv.bias, out, residual=xyz_normalized.contiguous(),
                        res_div=r)
        return out

    def forward(self, xyz, xyz_normalized, previous_level4=None, group=None, ragged=None, prev_point_major=False,
                **kwargs):
        """
        :param xyz Bx3xN input xyz, unnormalized; xyz_normalized Bx3xN; previous_level4 (Bx3xM, BxCxM) of the
               previous level (its batch may divide B: shared by consecutive patches)
        :param group (extension) patches per independent request, scope of group_knn's duplicate penalty
        :param prev_point_major (extension, eval) previous_level4[1] is laid out (clouds, M, C) instead of (B, C, M)
        :param ragged (extension, eval) operations.Ragged: request of every patch + valid size of every request's
               previous-level cloud, for requests with different numbers of patches
        :return xyz Bx3xNr (normalised frame), features BxCxN of the input points
        """
        if kwargs.get("phase") == "vis":
            raise NotImplementedError("phase='vis' (debug visualisation, upsampler.py:285-314) is out of scope")
        batch_size, _, num_point = xyz_normalized.size()
        fast = self._fast_path_ok(xyz_normalized)
        if ragged is not None and not fast:
            raise RuntimeError("ragged batches are an eval-mode (no-grad, CUDA fp32) feature")
        feat_pm_out = kwargs.pop("feat_pm_out", None)     # (extension, eval) buffer receiving the features point-major
        if fast and self.use_engine and self._engine_ok() and (previous_level4 is None or prev_point_major or self.fm_knn <= 0):
            return self._forward_engine(xyz, xyz_normalized, previous_level4, group, ragged, feat_pm_out)
        if feat_pm_out is not None:
            raise RuntimeError("feat_pm_out needs the level engine (eval mode, CUDA fp32, reference configuration)")
        if not fast and self.native_train and torch.is_grad_enabled() and not prev_point_major and ragged is None \
                and self._native_train_ok(xyz_normalized, previous_level4):
            prev_xyz, prev_feat = previous_level4 if (previous_level4 is not None and self.fm_knn > 0) else (None, None)
            return level_train.LevelTrainFunction.apply(self, xyz, xyz_normalized, prev_xyz, prev_feat, group,
                                                        *self._engine_params())
        if fast:
            x = self._features_fused(xyz_normalized, group, ragged)
        else:
            x = self.layer0(xyz_normalized.unsqueeze(dim=-1)).squeeze(dim=-1)
            y, _ = self.layer1(x)
            x = torch.cat([y, x], dim=1)
            y, _ = self.layer2(self.layer2_prep(x))
            x = torch.cat([y, x], dim=1)
            y, _ = self.layer3(self.layer3_prep(x))
            x = torch.cat([y, x], dim=1)
            y, _ = self.layer4(self.layer4_prep(x))
            x = torch.cat([y, x], dim=1)

        if previous_level4 is not None and self.fm_knn > 0:
            if prev_point_major:
                if not fast:
                    raise RuntimeError("point-major previous features are an eval-mode (no-grad, CUDA fp32) feature")
                x = self._skip_connection_fused(x, xyz, previous_level4, group, ragged)
            else:
                x = self._skip_connection(x, xyz, previous_level4, group, ragged)

        point_features = x
        if fast:
            return self._head_fused(x, xyz_normalized), point_features

        _, code_length, ratio = self.code.size()
        code = self._code_on(x.device).repeat(x.size(0), 1, num_point)
        x = x.unsqueeze(-1).expand(-1, -1, -1, ratio)
        x = torch.reshape(x, [batch_size, x.size(1), num_point * ratio]).contiguous()
        x = torch.cat([x, code], dim=1).unsqueeze(-1)
        x = self.up_layer(x)
        x = self.fc_layer1(x)
        x = self.fc_layer2(x).squeeze(-1)
        x = x + torch.reshape(xyz_normalized.unsqueeze(3).repeat([1, 1, 1, ratio]), [batch_size, 3, num_point * ratio])
        return x, point_features


class Net(torch.nn.Module):
    """3PU inter-level plus skip connection and dense layers (upsampler.py:9-189)."""

    def __init__(self, max_up_ratio=16, step_ratio=2, knn=16, growth_rate=12,
                 dense_n=3, max_num_point=312, fm_knn=3, **kwargs):
        super(Net, self).__init__()
        self.max_up_ratio = max_up_ratio
        self.step_ratio = step_ratio
        self.knn = knn
        self.growth_rate = growth_rate
        self.dense_n = dense_n
        self.fm_knn = fm_knn   # stored but, as in the reference (:25-26), not forwarded: Level's default 5 applies
        self.num_levels = int(log(max_up_ratio, step_ratio))
        self.levels = torch.nn.ModuleDict()
        self.max_num_point = max_num_point
        for l in range(1, self.num_levels + 1):
            self.levels['level_%d' % l] = Level(dense_n=dense_n, growth_rate=growth_rate, knn=knn, step_ratio=step_ratio)
        if self.training:
            for m in self.modules():
                if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv1d)):
                    torch.nn.init.xavier_uniform_(m.weight)
                    torch.nn.init.zeros_(m.bias)

    # ---- patch extraction (upsampler.py:39-105) ------------------------------------------------------
    def _train_patches(self, batch_xyz, k, gt_xyz, gt_k, seed_idx=None):
        batch_size, _, num_point = batch_xyz.size()
        if seed_idx is None:
            seed_idx = torch.randint(low=0, high=num_point, size=[batch_size, 1], dtype=torch.int32,
                                     device=batch_xyz.device)
        seeds = operations.gather_points(batch_xyz, seed_idx)                   # B x 3 x 1
        patches, _, _ = operations.group_knn(k, seeds, batch_xyz, unique=False, NCHW=True)
        patches = torch.cat(torch.unbind(patches, dim=2), dim=0)
        if gt_xyz is not None and gt_k is not None:
            gt_xyz, _, _ = operations.group_knn(gt_k, seeds, gt_xyz, unique=False)
            gt_xyz = torch.cat(torch.unbind(gt_xyz, dim=2), dim=0)
        else:
            gt_xyz = None
        return patches, gt_xyz

    def _eval_outlier_mask(self, batch_xyz):
        """:63-73: a point is kept when its nearest-neighbour distance is < 5x the cloud's mean."""
        _, _, closest_d = operations.group_knn(2, batch_xyz, batch_xyz, unique=False, NCHW=True)
        closest_d = closest_d[:, :, 1]
        return closest_d < (5 * torch.mean(closest_d, dim=1, keepdim=True))

    def _eval_tiles(self, batch_xyz, k):
        """Seeds by FPS, int(N/k*5) overlapping kNN tiles (:76-86) of ONE filtered cloud (1,3,N') -> (P,3,k'), P."""
        num_point = batch_xyz.shape[2]
        patch_num = int(num_point / k * 5)
        _, seeds = operations.furthest_point_sample(batch_xyz, patch_num)
        k = min(k, num_point)
        tiles, _, _ = operations.group_knn(k, seeds, batch_xyz, unique=False, NCHW=True)   # 1,3,P,k
        return torch.cat(torch.unbind(tiles, dim=2), dim=0), patch_num

    def extract_xyz_feature_patch(self, batch_xyz, k, gt_xyz=None, gt_k=None):
        """upsampler.py:39-105 (reference signature; eval expects batch 1 like the reference)."""
        if self.training:
            return self._train_patches(batch_xyz, k, gt_xyz, gt_k)
        assert batch_xyz.size(0) == 1
        mask = self._eval_outlier_mask(batch_xyz)
        batch_xyz = torch.masked_select(batch_xyz, mask.unsqueeze(1).expand_as(batch_xyz)).view(1, 3, -1)
        tiles, _ = self._eval_tiles(batch_xyz, k)
        return tiles, None

    # ---- eval levels past the first ------------------------------------------------------------------
    def _eval_level_single(self, level, xyz, old_xyz, old_features, max_num_point, num_output_point, **kwargs):
        """One request (batch 1) exactly as the reference walks it (upsampler.py:128-159); used when a filtered
        cloud is smaller than a tile, which changes the tile size itself."""
        if xyz.size(-1) > max_num_point:
            mask = self._eval_outlier_mask(xyz)
            xyz = torch.masked_select(xyz, mask.unsqueeze(1).expand_as(xyz)).view(1, 3, -1)
            patch_xyz, P = self._eval_tiles(xyz, max_num_point)
        else:
            patch_xyz, P = xyz, 1
        patch_norm, centroid, radius = operations.normalize_point_batch(patch_xyz, NCHW=True)
        new_xyz, features = level(patch_xyz, patch_norm, previous_level4=(old_xyz, old_features), group=P, **kwargs)
        new_xyz = new_xyz * radius + centroid
        if patch_xyz.shape[0] != 1:
            new_xyz = torch.cat(torch.split(new_xyz, 1, dim=0), dim=2)
            patch_xyz = torch.cat(torch.split(patch_xyz, 1, dim=0), dim=2)
            features = torch.cat(torch.split(features, 1, dim=0), dim=2)
            _, new_xyz = operations.furthest_point_sample(new_xyz, num_output_point)
        return new_xyz, patch_xyz, features

    def _eval_level_batched(self, level, xyz, old_xyz, old_features, old_n, max_num_point, num_output_point,
                            keep_features, **kwargs):
        """All requests of the batch together.  xyz (B,3,N) uniform; old_xyz (B,3,No) / old_features (B,264,No)
        padded, old_n (B,) int32 valid sizes.  The outlier filter leaves request b with N'_b points and
        int(N'_b/312*5) tiles, so tiles are processed as ONE flat list with an owner per tile.
        Returns (xyz (B,3,num_output_point), new old_xyz, new old_features, new old_n) or None when a filtered
        cloud is smaller than a tile (caller falls back to per-request processing)."""
        B, _, N = xyz.shape
        dev = xyz.device
        k = max_num_point
        mask = self._eval_outlier_mask(xyz)                                          # (B,N)
        counts = mask.sum(dim=1)
        counts_h = counts.tolist()                                                   # the one host sync of the level
        if min(counts_h) < k:
            return None
        # compact every cloud to its kept points, order preserved (what masked_select does, :72-73)
        order = torch.argsort((~mask).to(torch.uint8), dim=1, stable=True)
        xyz_c = torch.gather(xyz, 2, order.unsqueeze(1).expand(-1, 3, -1)).contiguous()
        n_arr = counts.to(torch.int32)
        P_h = [int(c / k * 5) for c in counts_h]                                     # :76
        Pmax = max(P_h)
        p_arr = torch.tensor(P_h, dtype=torch.int32, device=dev)
        req = torch.arange(B, dtype=torch.int32, device=dev)
        # seeds (:78) and tiles (:83)
        _, seeds = operations.furthest_point_sample_ragged(xyz_c, n_arr, p_arr, Pmax)  # (B,3,Pmax)
        tiles, _, _ = operations._knn_raw(k, seeds, xyz_c, False, None, want_dist=False,
                                          ragged=operations.Ragged(req, req, B, n_arr=n_arr, m_arr=p_arr))
        # flat list of the valid tiles, request-major (the reference's torch.cat(torch.unbind(., 2), 0), :85)
        owner_h = [b for b in range(B) for _ in range(P_h[b])]
        slot_h = [b * Pmax + p for b in range(B) for p in range(P_h[b])]
        owner = torch.tensor(owner_h, dtype=torch.int32, device=dev)
        slot = torch.tensor(slot_h, dtype=torch.int64, device=dev)
        patch_xyz = tiles.permute(0, 2, 1, 3).reshape(B * Pmax, 3, k)[slot].contiguous()   # (T,3,k)
        patch_norm, centroid, radius = operations.normalize_point_batch(patch_xyz, NCHW=True)
        ragged = operations.Ragged(owner, owner, B, n_arr=old_n)
        new_xyz, features = level(patch_xyz, patch_norm, previous_level4=(old_xyz, old_features), ragged=ragged,
                                  prev_point_major=True, **kwargs)
        new_xyz = new_xyz * radius + centroid                                        # (T,3,k*r)

        def merge(t):
            """(T,C,n) tiles -> (B,C,Pmax*n): the tiles of a request side by side (:149-155), zero padded"""
            C, n = t.shape[1], t.shape[2]
            buf = torch.zeros(B * Pmax, C, n, dtype=t.dtype, device=dev)
            buf[slot] = t
            return buf.view(B, Pmax, C, n).permute(0, 2, 1, 3).reshape(B, C, Pmax * n)

        merged = merge(new_xyz)
        r = new_xyz.shape[2] // k
        _, out_xyz = operations.furthest_point_sample_ragged(merged, p_arr * (k * r), None, num_output_point)  # :158
        if keep_features:
            # the next level gathers whole feature rows: hand the features over point-major, tiles side by side
            Cf = features.shape[1]
            feat_pm = torch.zeros(B, Pmax * k, Cf, dtype=torch.float32, device=dev)
            fused._lib.launch("pu3_to_point_major_f32", features, features.shape[0], Cf, k, features.data_ptr(),
                              slot.data_ptr(), feat_pm.data_ptr())
            return out_xyz, merge(patch_xyz), feat_pm, p_arr * k
        return out_xyz, None, None, None

    # Batched eval without host round trips.  The only data-dependent shapes of the eval path are the per-request tile
    # counts int(N'_b/312*5) after the outlier filter (:63-76).  N'_b <= N, so P = int(N/312*5) bounds them: every request
    # gets P tile SLOTS, the slots past its own count repeat its first tile (so that every kernel sees ordinary data and
    # the per-request duplicate-penalty scope is unchanged), and the counts stay on the device: they bound the seed FPS,
    # the merge FPS and the next level's skip search through the n_arr / m_arr arguments those kernels already take.
    # Nothing is read back, so the host enqueues the whole forward ahead of the GPU instead of stalling three times per
    # forward with ~25 small launches queued behind each stall (~1.7 ms of idle GPU per B=32 step).  The price is the
    # repeated tiles (none when the filter removes nothing, ~1/P of a level otherwise).  A filtered cloud smaller than one
    # tile changes the tile size itself: that is flagged on the device and the forward is redone on the synchronous path.
    static_tiles = True

    def _eval_level_static(self, level, xyz, old_xyz, old_features, old_n, k, num_output_point, keep_features, bad,
                           debug=None, **kwargs):
        """One level past the first for all requests, static shapes, every step a libpu3_b200 kernel (csrc/glue.cu for the
        steps the reference writes as torch expressions): outlier filter + compaction (:63-73), seed FPS (:78), kNN tiles
        (:83-85), normalisation (:138), Level, de-normalise + merge (:144-155), merge FPS (:158).  `bad` is an int32 device
        flag (a filtered cloud smaller than one tile)."""
        B, _, N = xyz.shape
        dev = xyz.device
        L = fused._lib
        r = self.step_ratio
        P = int(N / k * 5)                                                            # :76 with N' = N
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        key = (B, P, str(dev))
        cache = self.__dict__.setdefault("_static_cache", {})
        if key not in cache:
            cache[key] = (torch.arange(B, **i32), torch.arange(B, **i32).repeat_interleave(P))
        req, owner = cache[key]
        xyz = xyz.contiguous()
        # distance to the nearest other point (:63-66)
        _, _, closest = operations._knn_raw(2, xyz, xyz, False, None, want_knn=False)
        xyz_c, xyz_pm = torch.empty(B, 3, N, **f32), torch.empty(B, N, 3, **f32)
        n_arr, p_arr, pk_arr, pkr_arr = (torch.empty(B, **i32) for _ in range(4))
        L.launch("pu3_outlier_compact_f32", xyz, B, N, 2, k, r, closest.data_ptr(), xyz.data_ptr(), xyz_c.data_ptr(),
                 xyz_pm.data_ptr(), n_arr.data_ptr(), p_arr.data_ptr(), pk_arr.data_ptr(), pkr_arr.data_ptr(), bad.data_ptr())
        # seeds (:78): FPS over the kept points, p_arr[b] samples; spare tile slots repeat the request's first tile
        sidx = torch.zeros(B, P, **i32)
        L.launch("pu3_fps_ragged_f32", xyz_pm, B, N, P, n_arr.data_ptr(), p_arr.data_ptr(), xyz_pm.data_ptr(), None,
                 sidx.data_ptr(), tag="pu3_fps_f32")
        seeds = torch.empty(B, 3, P, **f32)
        L.launch("pu3_tile_seeds_f32", xyz_c, B, N, P, xyz_c.data_ptr(), sidx.data_ptr(), p_arr.data_ptr(), seeds.data_ptr())
        # tiles (:83-85) and their normalisation (:138)
        tiles, _, _ = operations._knn_raw(k, seeds, xyz_c, False, None, want_dist=False,
                                          ragged=operations.Ragged(req, req, B, n_arr=n_arr))      # (B,3,P,k)
        T = B * P
        patch_xyz, patch_norm = torch.empty(T, 3, k, **f32), torch.empty(T, 3, k, **f32)
        centroid, radius = torch.empty(T, 3, **f32), torch.empty(T, **f32)
        prev_xyz = torch.empty(B, 3, P * k, **f32) if keep_features else None
        L.launch("pu3_tiles_normalize_f32", tiles, B, P, k, tiles.data_ptr(), patch_xyz.data_ptr(), patch_norm.data_ptr(),
                 centroid.data_ptr(), radius.data_ptr(), L.ptr(prev_xyz))
        ragged = operations.Ragged(owner, owner, B, n_arr=old_n)
        feat_pm = None
        if keep_features and level.use_engine and level._engine_ok():
            feat_pm = torch.empty(B, P * k, 264, **f32)     # filled by the level itself (skip kernel / engine)
            kwargs = dict(kwargs, feat_pm_out=feat_pm)
        new_xyz, features = level(patch_xyz, patch_norm, previous_level4=(old_xyz, old_features), ragged=ragged,
                                  prev_point_major=True, **kwargs)                    # (T,3,k*r) normalised frame
        kr = new_xyz.shape[2]
        merged_pm = torch.empty(B, P * kr, 3, **f32)
        L.launch("pu3_denorm_merge_f32", new_xyz, B, P, kr, new_xyz.data_ptr(), centroid.data_ptr(), radius.data_ptr(),
                 merged_pm.data_ptr())
        if debug is not None:      # tests: intermediate state for stage-wise comparison with the oracle
            debug.update(patch_xyz=patch_xyz, merged_pm=merged_pm, p_arr=p_arr, n_arr=n_arr)
        # resample to num_output_point (:158)
        oidx = torch.zeros(B, num_output_point, **i32)
        L.launch("pu3_fps_ragged_f32", merged_pm, B, P * kr, num_output_point, pkr_arr.data_ptr(), None, merged_pm.data_ptr(),
                 None, oidx.data_ptr(), tag="pu3_fps_f32")
        out_xyz = torch.empty(B, 3, num_output_point, **f32)
        L.launch("pu3_gather_pm_f32", merged_pm, B, P * kr, num_output_point, merged_pm.data_ptr(), oidx.data_ptr(),
                 out_xyz.data_ptr())
        if keep_features:
            if feat_pm is None:
                Cf = features.shape[1]
                feat_pm = torch.empty(B, P * k, Cf, **f32)
                L.launch("pu3_to_point_major_f32", features, features.shape[0], Cf, k, features.data_ptr(), None, feat_pm.data_ptr())
            return out_xyz, prev_xyz, feat_pm, pk_arr
        return out_xyz, None, None, None

    def forward(self, xyz, ratio=None, gt=None, seed_idx_per_level=None, **kwargs):
        """
        :param xyz Bx3xN; ratio upscaling factor; gt Bx3x(max_up_ratio*N) (training)
        :param seed_idx_per_level (extension, training) {level: (B,1) int32} to fix the random zoom seeds
        :return xyz Bx3x(ratio*N) (eval) or (xyz, gt) zoomed patches (training)
        """
        ratio = ratio or self.max_up_ratio
        if self.training:
            assert gt is not None
        batch_size, _, num_point = xyz.size()
        num_levels = int(log(ratio, self.step_ratio))
        max_num_point = min(num_point, self.max_num_point)
        if not self.training:
            return self._forward_eval(xyz, num_levels, num_point, max_num_point, **kwargs)

        for l in range(1, num_levels + 1):
            curr_ratio = self.step_ratio ** l
            level = self.levels['level_%d' % l]
            if l == 1:
                old_xyz = xyz
                xyz, features = level(xyz, xyz, previous_level4=None, **kwargs)
                old_features = features
                continue
            if xyz.size(-1) > max_num_point:
                gt_k = max_num_point * ratio // curr_ratio * self.step_ratio
                sidx = None if seed_idx_per_level is None else seed_idx_per_level.get(l)
                patch_xyz, gt = self._train_patches(xyz, max_num_point, gt, gt_k, seed_idx=sidx)
            else:
                patch_xyz = xyz
            patch_norm, centroid, radius = operations.normalize_point_batch(patch_xyz, NCHW=True)
            xyz, features = level(patch_xyz, patch_norm, previous_level4=(old_xyz, old_features), **kwargs)
            xyz = xyz * radius + centroid
            old_xyz, old_features = patch_xyz, features
        return xyz, gt

    # Request groups processed concurrently on separate CUDA streams (eval).  FPS is a chain of dependent
    # rounds that leaves most of the chip idle (cluster hand-shake latency); a second group's kNN / MLP kernels
    # can fill those SMs.  Measured on B200 (profiles/r1_bench_history.md): no gain at B=32 -- the FPS kernels
    # already occupy 128 of the 148 SMs with register-heavy CTAs -- so the default is one group.
    eval_groups = 1

    def _forward_eval(self, xyz, num_levels, num_point, max_num_point, **kwargs):
        """Eval: every cloud of the batch is an independent request (the reference takes one per call)."""
        B = xyz.shape[0]
        groups = self.eval_groups or 1
        groups = max(1, min(int(groups), B))
        if groups == 1 or num_levels < 2:
            return self._forward_eval_group(xyz, num_levels, num_point, max_num_point, **kwargs)
        # one host thread + one stream per group: the per-level host reads (outlier counts) of one group must not
        # stall the launches of the other
        import threading
        cur = torch.cuda.current_stream(xyz.device)
        bounds = [(g * B) // groups for g in range(groups + 1)]
        outs, errs = [None] * groups, [None] * groups
        streams = self._eval_streams(xyz.device, groups)
        grad = torch.is_grad_enabled()

        def work(g):
            try:
                with torch.cuda.device(xyz.device), torch.cuda.stream(streams[g]), torch.set_grad_enabled(grad):
                    streams[g].wait_stream(cur)
                    outs[g] = self._forward_eval_group(xyz[bounds[g]:bounds[g + 1]], num_levels, num_point, max_num_point, **kwargs)
            except BaseException as e:   # re-raised in the caller's thread
                errs[g] = e

        threads = [threading.Thread(target=work, args=(g,)) for g in range(1, groups)]
        for t in threads:
            t.start()
        work(0)
        for t in threads:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        for g in range(groups):
            cur.wait_stream(streams[g])
            outs[g].record_stream(cur)
        return torch.cat(outs, dim=0)

    def _eval_streams(self, device, n):
        cache = self.__dict__.setdefault("_stream_cache", {})
        key = (device, n)
        if key not in cache:
            cache[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
        return cache[key]

    # The static forward is a fixed sequence of ~230 launches on fixed shapes: captured once per (shape, parameter storage) in
    # a CUDA graph and replayed (no Python / ctypes / allocator work per launch, no gaps between the small kernels).
    use_cuda_graph = True

    def _forward_eval_group(self, xyz, num_levels, num_point, max_num_point, **kwargs):
        if self.static_tiles and xyz.is_cuda and num_levels > 1 and xyz.shape[2] >= max_num_point and not kwargs:
            prof = fused._lib._profiler
            if self.use_cuda_graph and xyz.dtype == torch.float32 and not (prof is not None and prof.timing):
                out, bad = self._forward_eval_graph(xyz, num_levels, num_point, max_num_point)
            else:
                out, bad = self._forward_eval_static(xyz, num_levels, num_point, max_num_point)
            if out is not None and not bool(bad):     # the one host read of the forward, after everything is enqueued
                return out
        return self._forward_eval_sync(xyz, num_levels, num_point, max_num_point, **kwargs)

    def _forward_eval_static(self, xyz, num_levels, num_point, max_num_point):
        """Every level with static shapes (see static_tiles above).  Returns (xyz, bad flag on the device) or (None, None)
        when the shapes do not tile."""
        B = xyz.shape[0]
        dev = xyz.device
        xyz = xyz.contiguous()
        old_xyz = xyz
        xyz, feats = self.levels['level_1'](xyz, xyz, previous_level4=None, group=1)
        old_n = torch.full((B,), old_xyz.shape[2], dtype=torch.int32, device=dev)
        pm = torch.empty(B, feats.shape[2], feats.shape[1], dtype=torch.float32, device=dev)
        fused._lib.launch("pu3_to_point_major_f32", feats, B, feats.shape[1], feats.shape[2], feats.contiguous().data_ptr(),
                          None, pm.data_ptr())
        old_features = pm
        bad = torch.zeros((), dtype=torch.int32, device=dev)
        for l in range(2, num_levels + 1):
            if xyz.size(-1) <= max_num_point:
                return None, None
            xyz, old_xyz, old_features, old_n = self._eval_level_static(
                self.levels['level_%d' % l], xyz, old_xyz, old_features, old_n, max_num_point,
                num_point * self.step_ratio ** l, l < num_levels, bad)
        return xyz, bad

    def _graph_key(self, xyz, num_levels):
        ptrs = tuple(p.data_ptr() for p in self.parameters())
        return (tuple(xyz.shape), num_levels, str(xyz.device), hash(ptrs))

    def _forward_eval_graph(self, xyz, num_levels, num_point, max_num_point):
        """_forward_eval_static through a captured CUDA graph.  The graph bakes in the addresses of the parameters (their
        VALUES are read at replay: weight updates are seen) and of its private input / output buffers; it is re-captured
        when the input shape or the parameter storage changes."""
        cache = self.__dict__.setdefault("_graph_cache", {})
        key = self._graph_key(xyz, num_levels)
        entry = cache.get(key)
        if entry is None:
            if len(cache) >= 4:                       # bounded: every graph keeps its workspace pool alive
                cache.pop(next(iter(cache)))
            cur = torch.cuda.current_stream(xyz.device)
            side = torch.cuda.Stream(device=xyz.device)
            static_in = torch.empty_like(xyz, memory_format=torch.contiguous_format)
            static_in.copy_(xyz)
            side.wait_stream(cur)
            with torch.cuda.stream(side):             # eager warm-up: lazy initialisation (function attributes, caches)
                out, _ = self._forward_eval_static(static_in, num_levels, num_point, max_num_point)
            cur.wait_stream(side)
            if out is None:
                cache[key] = False
                return None, None
            torch.cuda.synchronize(xyz.device)
            graph = torch.cuda.CUDAGraph()
            counter = fused._lib.Profiler(timing=False)
            prev = fused._lib.set_profiler(counter)
            try:
                with torch.cuda.graph(graph, stream=side):
                    out, bad = self._forward_eval_static(static_in, num_levels, num_point, max_num_point)
            finally:
                fused._lib.set_profiler(prev)
            entry = cache[key] = (graph, static_in, out, bad, counter.launches)
        if entry is False:
            return None, None
        graph, static_in, out, bad, launches = entry
        static_in.copy_(xyz)
        graph.replay()
        prof = fused._lib._profiler
        if prof is not None:
            prof.launches += launches
            prof.calls["cuda_graph[eval forward]"] = prof.calls.get("cuda_graph[eval forward]", 0) + 1
        return out.clone(), bad

    def _forward_eval_sync(self, xyz, num_levels, num_point, max_num_point, **kwargs):
        B = xyz.shape[0]
        dev = xyz.device
        level = self.levels['level_1']
        old_xyz = xyz
        xyz, old_features = level(xyz, xyz, previous_level4=None, group=1, **kwargs)    # duplicate-penalty scope: 1 cloud
        old_n = torch.full((B,), old_xyz.shape[2], dtype=torch.int32, device=dev)
        point_major = xyz.is_cuda and num_levels > 1
        if point_major:    # (B,C,N) -> (B,N,C): the layout the fused skip connection gathers from
            pm = torch.empty(B, old_features.shape[2], old_features.shape[1], dtype=torch.float32, device=dev)
            fused._lib.launch("pu3_to_point_major_f32", old_features, B, old_features.shape[1], old_features.shape[2],
                              old_features.contiguous().data_ptr(), None, pm.data_ptr())
            old_features = pm
        for l in range(2, num_levels + 1):
            level = self.levels['level_%d' % l]
            num_output_point = num_point * self.step_ratio ** l
            res = None
            if xyz.size(-1) > max_num_point and xyz.is_cuda:
                res = self._eval_level_batched(level, xyz, old_xyz, old_features, old_n, max_num_point, num_output_point,
                                               keep_features=l < num_levels, **kwargs)
            if res is not None:
                xyz, old_xyz, old_features, old_n = res
                continue
            # per-request processing (a filtered cloud smaller than one tile, or nothing to tile)
            old_n_h = old_n.tolist()
            if point_major:
                old_features, point_major = old_features.transpose(1, 2), False    # (B,C,M) view for this path
            outs = [self._eval_level_single(level, xyz[i:i + 1], old_xyz[i:i + 1, :, :old_n_h[i]].contiguous(),
                                            old_features[i:i + 1, :, :old_n_h[i]].contiguous(), max_num_point,
                                            num_output_point, **kwargs) for i in range(B)]
            xyz = torch.cat([o[0] for o in outs], dim=0)
            old_n = torch.tensor([o[1].shape[2] for o in outs], dtype=torch.int32, device=dev)
            nmax = int(old_n.max())
            pad = lambda t: torch.nn.functional.pad(t, (0, nmax - t.shape[2]))
            old_xyz = torch.cat([pad(o[1]) for o in outs], dim=0)
            old_features = torch.cat([pad(o[2]) for o in outs], dim=0)
            if xyz.is_cuda and num_levels > 1:   # restore the invariant: previous features travel point-major
                old_features, point_major = old_features.transpose(1, 2).contiguous(), True
        return xyz
