"""VAE^2 HRNet models on the B200 engine  --  drop-in for reference lib/models/enc_hrnet.py.

Same factories (``get_encdec_model / get_encz_model / get_D_sequence_model /
get_D_frame_model``, reference :1185-1210), same class names, constructor arguments, forward
signatures, attributes (``hd_z``, ``z_dim``, ``clip_length``) and -- because the parameter
tree is registered under the reference's attribute names -- the same ``state_dict`` keys and
shapes, so reference checkpoints and the ImageNet-pretrained remap (:761-785) load unchanged.

What differs is execution: the nn.Conv2d / BatchNorm2d children are parameter containers
only.  ``forward`` records the whole network once per input shape into a static launch plan
(engine/graph.py) of hand-written sm_100a kernels and replays it as a single autograd node.
"""
import logging
import os

import numpy as np
import torch
import torch.nn as nn

from ._engine import EngineModule

BatchNorm2d = nn.BatchNorm2d
BN_MOMENTUM = 0.01          # reference :23
logger = logging.getLogger(__name__)


def _c(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=bias)


def _b(c):
    return BatchNorm2d(c, momentum=BN_MOMENTUM)


def conv3x3(in_planes, out_planes, stride=1):
    """3x3 convolution with padding (reference :27-30)."""
    return _c(in_planes, out_planes, 3, stride)


def _cbr(cin, cout, k, stride, relu):
    layers = [_c(cin, cout, k, stride), _b(cout)]
    if relu:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class _Single(EngineModule):
    """Standalone execution of a one-tensor-in / one-tensor-out module that defines ``emit``."""

    def _record(self, rec, shapes, needs, tag):
        (_, C_, H, W), = shapes
        y = self.emit(rec, rec.input(C_, H, W, needs[0]))
        rec.output(y)
        return [(y.C, y.H, y.W)]

    def forward(self, x):
        return self._run([x])[0]


class BasicBlock(_Single):
    """conv3x3-BN-ReLU-conv3x3-BN-(+x)-ReLU   (reference :33-62)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = conv3x3(inplanes, planes, stride), _b(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2, self.bn2 = conv3x3(planes, planes), _b(planes)
        self.downsample, self.stride = downsample, stride

    def emit(self, rec, x, out=None):
        return self.emit_multi(rec, [self], [x], [out])[0]

    @staticmethod
    def emit_multi(rec, blocks, xs, outs):
        """The same block position of several independent branches, in lock-step: conv1 of every branch,
        ONE BN group, conv2 of every branch, ONE BN group (so SyncBN needs one collective per depth)."""
        hs = rec.conv_bn_multi([dict(x=x, conv=b.conv1, bn=b.bn1, relu=True) for b, x in zip(blocks, xs)])
        ress = [x if b.downsample is None else rec.conv_bn(x, b.downsample[0], b.downsample[1])
                for b, x in zip(blocks, xs)]
        return rec.conv_bn_multi([dict(x=h, conv=b.conv2, bn=b.bn2, relu=True, residual=r, out=o)
                                  for b, h, r, o in zip(blocks, hs, ress, outs)])


class Bottleneck(_Single):
    """1x1-BN-ReLU-3x3-BN-ReLU-1x1(x4)-BN-(+res)-ReLU   (reference :65-103)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = _c(inplanes, planes, 1), _b(planes)
        self.conv2, self.bn2 = _c(planes, planes, 3, stride), _b(planes)
        self.conv3, self.bn3 = _c(planes, planes * self.expansion, 1), _b(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample, self.stride = downsample, stride

    def emit(self, rec, x, out=None):
        h = rec.conv_bn(x, self.conv1, self.bn1, relu=True)
        h = rec.conv_bn(h, self.conv2, self.bn2, relu=True)
        res = x if self.downsample is None else rec.conv_bn(x, self.downsample[0], self.downsample[1])
        return rec.conv_bn(h, self.conv3, self.bn3, relu=True, residual=res, out=out)


blocks_dict = {"BASIC": BasicBlock, "BOTTLENECK": Bottleneck}


def _block_stack(block, inplanes, planes, n, stride=1):
    """``_make_layer`` / ``_make_one_branch`` (reference :144-166, :408-423)."""
    down = None
    if stride != 1 or inplanes != planes * block.expansion:
        down = nn.Sequential(_c(inplanes, planes * block.expansion, 1, stride), _b(planes * block.expansion))
    layers = [block(inplanes, planes, stride, down)]
    layers += [block(planes * block.expansion, planes) for _ in range(1, n)]
    return nn.Sequential(*layers)


def _emit_stack(rec, stack, x, out=None):
    n = len(stack)
    for i, blk in enumerate(stack):
        x = blk.emit(rec, x, out=out if i == n - 1 else None)
    return x


def _emit_stacks_lockstep(rec, stacks, xs, outs=None):
    """Independent block stacks (the branches of one HR module) advanced block by block."""
    if not all(isinstance(b, BasicBlock) for st in stacks for b in st):
        return [_emit_stack(rec, st, x, outs[i] if outs else None) for i, (st, x) in enumerate(zip(stacks, xs))]
    cur = list(xs)
    for k in range(max(len(st) for st in stacks)):
        act = [b for b in range(len(stacks)) if k < len(stacks[b])]
        o = [outs[b] if (outs and k == len(stacks[b]) - 1) else None for b in act]
        res = BasicBlock.emit_multi(rec, [stacks[b][k] for b in act], [cur[b] for b in act], o)
        for b, r in zip(act, res):
            cur[b] = r
    return cur


class HighResolutionModule(EngineModule):
    """Parallel branches + multi-resolution fusion (reference :106-250)."""

    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method,
                 multi_scale_output=True):
        super().__init__()
        for what, lst in (("NUM_BLOCKS", num_blocks), ("NUM_CHANNELS", num_channels),
                          ("NUM_INCHANNELS", num_inchannels)):
            if num_branches != len(lst):
                msg = "NUM_BRANCHES({}) <> {}({})".format(num_branches, what, len(lst))
                logger.error(msg)
                raise ValueError(msg)
        self.num_inchannels, self.fuse_method = num_inchannels, fuse_method
        self.num_branches, self.multi_scale_output = num_branches, multi_scale_output
        branches = []
        for i in range(num_branches):
            branches.append(_block_stack(blocks, num_inchannels[i], num_channels[i], num_blocks[i]))
            self.num_inchannels[i] = num_channels[i] * blocks.expansion
        self.branches = nn.ModuleList(branches)
        self.fuse_layers = self._make_fuse_layers()
        self.relu = nn.ReLU(inplace=True)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return None
        ch = self.num_inchannels
        rows = []
        for i in range(self.num_branches if self.multi_scale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j > i:        # lower resolution -> 1x1 conv + BN, then bilinear up (in forward)
                    row.append(_cbr(ch[j], ch[i], 1, 1, relu=False))
                elif j == i:
                    row.append(None)
                else:            # higher resolution -> chain of (i-j) stride-2 3x3 convs
                    steps = [_cbr(ch[j], ch[i] if k == i - j - 1 else ch[j], 3, 2, relu=(k != i - j - 1))
                             for k in range(i - j)]
                    row.append(nn.Sequential(*steps))
            rows.append(nn.ModuleList(row))
        return nn.ModuleList(rows)

    def get_num_inchannels(self):
        return self.num_inchannels

    def emit(self, rec, xs, outs=None):
        """xs: list of branch activations -> list of fused activations (reference :226-250).
        ``outs[i]`` optionally names the buffer (e.g. a concat slice) output i must land in."""
        nb = self.num_branches
        if nb == 1:
            return [_emit_stack(rec, self.branches[0], xs[0], out=outs[0] if outs else None)]
        xs = _emit_stacks_lockstep(rec, list(self.branches), xs)
        # fuse convs level by level: level 0 = every 1x1 (j > i) and the first stride-2 step of every
        # down chain (j < i), level 1 = second steps, ... -- each level is one BN group
        n_out = len(self.fuse_layers)
        terms = [[None] * nb for _ in range(n_out)]
        cur = {}
        for i in range(n_out):
            for j in range(nb):
                if j == i:
                    terms[i][j] = xs[j]
                elif j < i:
                    cur[(i, j)] = xs[j]
        level = 0
        while True:
            items, keys = [], []
            for i in range(n_out):
                for j in range(nb):
                    if j > i and level == 0:
                        fl = self.fuse_layers[i][j]
                        items.append(dict(x=xs[j], conv=fl[0], bn=fl[1], relu=False))
                        keys.append((i, j, True))
                    elif j < i and level < len(self.fuse_layers[i][j]):
                        step = self.fuse_layers[i][j][level]
                        items.append(dict(x=cur[(i, j)], conv=step[0], bn=step[1], relu=len(step) == 3))
                        keys.append((i, j, level == len(self.fuse_layers[i][j]) - 1))
            if not items:
                break
            for (i, j, last), t in zip(keys, rec.conv_bn_multi(items)):
                if last:
                    terms[i][j] = t
                else:
                    cur[(i, j)] = t
            level += 1
        # reference sums j = 0..nb-1 in order; fp32 addition order is kept (terms[i][0] first)
        return [rec.fuse(terms[i], xs[i].H, xs[i].W, relu=True, out=outs[i] if outs else None) for i in range(n_out)]

    def _record(self, rec, shapes, needs, tag):
        xs = [rec.input(s[1], s[2], s[3], n) for s, n in zip(shapes, needs)]
        ys = self.emit(rec, xs)
        for y in ys:
            rec.output(y)
        return [(y.C, y.H, y.W) for y in ys]

    def forward(self, x):
        return list(self._run(list(x)))


def _transition(pre, cur):
    """``_make_transition_layer`` (reference :372-406)."""
    layers = []
    for i, c in enumerate(cur):
        if i < len(pre):
            layers.append(_cbr(pre[i], c, 3, 1, relu=True) if c != pre[i] else None)
        else:
            n = i + 1 - len(pre)
            layers.append(nn.Sequential(*[_cbr(pre[-1], c if j == n - 1 else pre[-1], 3, 2, relu=True)
                                          for j in range(n)]))
    return nn.ModuleList(layers)


def _emit_transition(rec, layers, prev, n_pre, outs=None):
    """Existing branches pass through (or get a 3x3 conv when widths differ); new branches are
    stride-2 chains fed from the LAST previous branch (reference forward :796-817)."""
    res = [None] * len(layers)
    items, idx, chains = [], [], {}
    for i, layer in enumerate(layers):
        out = outs[i] if outs else None
        if i < n_pre:
            if layer is None:
                res[i] = prev[i] if out is None else rec.copy(prev[i], out)
            else:
                items.append(dict(x=prev[i], conv=layer[0], bn=layer[1], relu=True, out=out))
                idx.append(i)
        else:
            chains[i] = prev[-1]
    level = 0
    while True:      # new branches: stride-2 chains from the last previous branch, advanced level by level
        for i, t in chains.items():
            layer = layers[i]
            if level < len(layer):
                last = level == len(layer) - 1
                items.append(dict(x=t, conv=layer[level][0], bn=layer[level][1], relu=True,
                                  out=(outs[i] if outs else None) if last else None))
                idx.append(i)
        if not items:
            break
        for i, t in zip(idx, rec.conv_bn_multi(items)):
            if i < n_pre:
                res[i] = t
            else:
                chains[i] = t
                res[i] = t
        items, idx = [], []
        level += 1
    return res


def _head(cin, cout, k):
    """last_layer*: 1x1(+bias) - BN - ReLU - kxk(+bias)   (reference :323-338)."""
    return nn.Sequential(_c(cin, cin, 1, bias=True), _b(cin), nn.ReLU(inplace=True), _c(cin, cout, k, bias=True))


def _stage_channels(cfg):
    exp = blocks_dict[cfg["BLOCK"]].expansion
    return [c * exp for c in cfg["NUM_CHANNELS"]]


class HighResolutionNet(EngineModule):
    """Shared trunk builder.  ``prefixes`` lists the attribute prefixes of the HRNets this module
    owns ('' for the encoder / posterior / discriminator, 'decf_' and 'decp_' for the decoders)."""

    def __init__(self, config, in_channels=None, prefixes=("",), code_extra=None, heads=3, **kwargs):
        super().__init__()
        extra = config.MODEL.EXTRA
        self.is_baseline = extra.IS_BASELINE
        self.enable_random_code = kwargs.get("enable_random_code", False)
        self.clip_length = config.TRAIN.CLIP_LENGTH
        self.hd_z, self.z_dim = extra.HD_Z, extra.Z_DIM
        self.extra = extra
        self.stage1_cfg, self.stage2_cfg = extra["STAGE1"], extra["STAGE2"]
        self.stage3_cfg, self.stage4_cfg = extra["STAGE3"], extra["STAGE4"]
        in_channels = in_channels or {"": 3}
        code_extra = code_extra or {}
        for p in prefixes:
            last = self._build_trunk(p, in_channels[p], code_extra.get(p))
            if heads and p == "":
                self.last_stage_channels = last
                self.last_inp_channels = int(np.sum(last))
            for h in range(1, heads + 1):
                setattr(self, "%slast_layer_%d" % (p, h),
                        _head(int(np.sum(last)), config.DATASET.NUM_CLASSES, extra.FINAL_CONV_KERNEL))
        if not heads:
            self.last_stage_channels = last
            self.last_inp_channels = int(np.sum(last))
            self.last_layer_1 = self.last_layer_2 = self.last_layer_3 = None

    def _build_trunk(self, p, cin, code_ch):
        s1 = self.stage1_cfg
        setattr(self, p + "conv1", _c(cin, 64, 3, 1))
        setattr(self, p + "bn1", _b(64))
        setattr(self, p + "conv2", _c(64, 64, 3, 1))
        setattr(self, p + "bn2", _b(64))
        setattr(self, p + "relu", nn.ReLU(inplace=True))
        blk = blocks_dict[s1["BLOCK"]]
        setattr(self, p + "layer1", _block_stack(blk, 64, s1["NUM_CHANNELS"][0], s1["NUM_BLOCKS"][0]))
        pre = [blk.expansion * s1["NUM_CHANNELS"][0]]
        for n, cfg in ((2, self.stage2_cfg), (3, self.stage3_cfg), (4, self.stage4_cfg)):
            ch = _stage_channels(cfg)
            setattr(self, "%stransition%d" % (p, n - 1), _transition(pre, ch))
            if n == 4 and code_ch:
                setattr(self, p + "transition3_e", _transition([c + code_ch for c in ch], ch))
            mods = []
            for _ in range(cfg["NUM_MODULES"]):
                mods.append(HighResolutionModule(cfg["NUM_BRANCHES"], blocks_dict[cfg["BLOCK"]], cfg["NUM_BLOCKS"],
                                                 ch, cfg["NUM_CHANNELS"], cfg["FUSE_METHOD"], True))
                ch = mods[-1].get_num_inchannels()
            setattr(self, "%sstage%d" % (p, n), nn.Sequential(*mods))
            pre = ch
        return pre

    # ---- recording -----------------------------------------------------------------------
    def _emit_trunk(self, rec, p, x, code_maps=None, head_cat=True, tile=1):
        """stem -> layer1 -> stage2..4.  ``code_maps(b, H, W, C)`` (optional) returns the concat
        root and feature slice for branch b in front of transition3_e.  Returns the stage-4
        outputs; with head_cat they are laid out inside the head's concat buffer.
        ``tile`` = K > 1 (K-sample inference): x holds B context clips, everything up to transition3 runs once per
        clip and its result is replicated K times into the concat buffers, which (like all that follows) hold K*B
        samples -- the K latent draws of every clip."""
        g = lambda n: getattr(self, p + n)
        x = rec.conv_bn(x, g("conv1"), g("bn1"), relu=True)
        x = rec.conv_bn(x, g("conv2"), g("bn2"), relu=True)
        taps = rec.plan.taps          # module-boundary activations by the oracle's tap names (parity tests)
        taps[p + "stem"] = x
        x = _emit_stack(rec, g("layer1"), x)
        taps[p + "layer1"] = x
        xs = _emit_transition(rec, g("transition1"), [x], 1)
        ys = self._emit_stage(rec, g("stage2"), xs)
        xs = _emit_transition(rec, g("transition2"), ys, len(ys))
        ys = self._emit_stage(rec, g("stage3"), xs)
        for i, t in enumerate(ys):
            taps["%sstage3.%d" % (p, i)] = t
        if code_maps is None:
            xs = _emit_transition(rec, g("transition3"), ys, len(ys))
        else:
            t3 = g("transition3")
            sizes = self._branch_sizes(ys, len(t3))
            cats = [code_maps(b, h, w, c) for b, (c, h, w) in enumerate(sizes)]
            if tile > 1:
                feats = _emit_transition(rec, t3, ys, len(ys))
                for f, (_, dst) in zip(feats, cats):
                    rec.tile(f, dst, tile)
            else:
                _emit_transition(rec, t3, ys, len(ys), outs=[feat for _, feat in cats])
            xs = _emit_transition(rec, g("transition3_e"), [root for root, _ in cats], len(cats))
        outs = None
        if head_cat:
            cat, slices = rec.concat([a.C for a in xs], xs[0].H, xs[0].W, name=p + "headcat", B=xs[0].B)
            outs = [slices[0]] + [None] * (len(xs) - 1)
        ys = self._emit_stage(rec, g("stage4"), xs, outs)
        for i, t in enumerate(ys):
            taps["%sstage4.%d" % (p, i)] = t
        if not head_cat:
            return ys, None
        for b in range(1, len(ys)):      # bilinear up-sample straight into the concat slices (:833-839)
            rec.fuse([ys[b]], cat.H, cat.W, relu=False, out=slices[b])
        return ys, cat

    @staticmethod
    def _emit_stage(rec, stage, xs, outs=None):
        n = len(stage)
        for m, mod in enumerate(stage):
            xs = mod.emit(rec, xs, outs if m == n - 1 else None)
        return xs

    def _branch_sizes(self, ys, n):
        ch = _stage_channels(self.stage4_cfg)
        sizes = [(ch[i], ys[i].H, ys[i].W) for i in range(len(ys))]
        h, w = ys[-1].H, ys[-1].W
        for i in range(len(ys), n):
            h, w = (h + 1) // 2, (w + 1) // 2
            sizes.append((ch[i], h, w))
        return sizes

    @staticmethod
    def _emit_head(rec, head, cat, out=None):
        h = rec.conv_bn(cat, head[0], head[1], relu=True)
        return rec.conv(h, head[3], y=out)

    def init_weights(self, pretrained=""):
        """Reference init (:753-760): conv N(0, 1e-3), BN (1, 0); then the optional pretrained remap."""
        logger.info("=> init weights from normal distribution")
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if pretrained and os.path.isfile(pretrained):
            self._load_pretrained(torch.load(pretrained, map_location="cpu"), pretrained)

    def _conv1_repeat(self):
        return None

    def _load_pretrained(self, blob, path):
        """ImageNet HRNet weights -> this module (reference :761-785, :1051-1068, :1164-1183):
        strip 'model.', skip heads, tile conv1 over the clip's frames, copy into decoders."""
        own = self.state_dict()
        picked = {k.replace("model.", ""): v for k, v in blob.items()
                  if k.replace("model.", "") in own and "last_layer" not in k}
        rep = self._conv1_repeat()
        extra = {}
        for k, v in picked.items():
            if k == "conv1.weight" and rep is not None:
                extra[k] = v.repeat([1, rep[""], 1, 1])
            for p in ("decf_", "decp_"):
                if hasattr(self, p + "conv1"):
                    extra[p + k] = v.repeat([1, rep[p], 1, 1]) if (k == "conv1.weight" and rep) else v
        picked.update(extra)
        logger.info("=> loading %d tensors from pretrained model %s", len(picked), path)
        own.update(picked)
        self.load_state_dict(own)


class HighResolutionNetED(HighResolutionNet):
    """Encoder + future/past decoders (reference :530-981)."""

    def __init__(self, config, **kwargs):
        extra = config.MODEL.EXTRA
        L = config.TRAIN.CLIP_LENGTH
        coded = extra.BASELINE_MODE != "DETERMINISTIC"
        z = extra.Z_DIM
        super().__init__(
            config, prefixes=("", "decf_", "decp_"),
            in_channels={"": 3 * L * 2 if extra.IS_BASELINE else 3 * L, "decf_": 3 * L, "decp_": 3 * L},
            code_extra={"": (z if extra.IS_BASELINE else 2 * z) if coded else 0,
                        "decf_": z if coded else 0, "decp_": z if coded else 0},
            heads=3, enable_random_code=coded, **kwargs)

    def _conv1_repeat(self):
        L = self.clip_length
        return {"": L * 2 if self.extra.IS_BASELINE else L, "decf_": L, "decp_": L}

    def _record(self, rec, shapes, needs, tag):
        B, Cx, H, W = shapes[0]
        Z, coded = self.z_dim, self.enable_random_code
        K = tag[1] if isinstance(tag, tuple) and tag[0] == "ksample" else 1
        KB = K * B                                  # plan.B; the context clips (and the encoder trunk) have B samples
        x = rec.input(Cx, H, W, needs[0], B=B)
        nz = 4 if self.hd_z else 1            # HD_Z: one latent map per branch; otherwise ONE per-sample z [B, Z, 1, 1]
        z_slots = [rec.new_input_slot() for _ in range(nz)] if coded else []
        code_slot = rec.new_input_slot() if (coded and not self.is_baseline) else None
        z_needs = needs[1:1 + nz] if coded else []

        def maps_for(first):
            pending = {}

            def code_maps(b, h, w, c):
                segs = ([Z] if (first and not self.is_baseline) else []) + [Z, c]
                root, sl = rec.concat(segs, h, w, name="zcat%d" % b, B=KB)
                if len(segs) == 3:
                    rec.code(code_slot, Z, sl[0])
                pending[b] = sl[-2]
                return root, sl[-1]
            return code_maps, pending

        z_dsts = [[] for _ in range(4)]
        preds = []
        cur = x
        for p, first in (("", True), ("decf_", False), ("decp_", False)):
            cm, pending = maps_for(first) if coded else (None, {})
            _, cat = self._emit_trunk(rec, p, cur if first else preds[0][0], cm, tile=K if first else 1)
            for b, sl in pending.items():
                z_dsts[b].append(sl)
            root, sl = rec.concat([3, 3, 3], H, W, name=p + "pred", B=KB)
            heads = [getattr(self, "%slast_layer_%d" % (p, h + 1)) for h in range(3)]
            mids = rec.conv_bn_multi([dict(x=cat, conv=hd[0], bn=hd[1], relu=True) for hd in heads])
            for h in range(3):
                rec.conv(mids[h], heads[h][3], y=sl[h])
            preds.append((root, sl))
        # z inputs fan out into the three nets' concat buffers (enc_hrnet.py:825-826, 885, 943)
        for b in range(4 if coded else 0):
            if self.hd_z:
                _, zc, zh, zw = shapes[1 + b]
                rec.input(zc, zh, zw, z_needs[b], slot=z_slots[b], into=z_dsts[b])
            else:                              # _gen_code_map(x_list, z): the per-sample z repeated over every map (:819)
                for dst in z_dsts[b]:
                    rec.code(z_slots[0], Z, dst, needs_grad=z_needs[0])
        # outputs in the reference's return order: (x1t, x2t, x3t) = (decp, enc, decf)   (:981)
        for root, sl in (preds[2], preds[0], preds[1]):
            slot = rec.new_output_slot()
            for h in range(3):
                rec.output(sl[h], dst_ctot=9, dst_coff=3 * h, slot=slot)
        return [(9, H, W)] * 3

    def forward(self, x, z=None, is_baseline=False, *args, **kwargs):
        if is_baseline or self.is_baseline:
            raise NotImplementedError("vae2_b200: the IS_BASELINE ablation path is not built (SURVEY.md §8 scope)")
        inputs = [x]
        if self.enable_random_code:
            if self.hd_z:
                if not (isinstance(z, (list, tuple)) and len(z) == 4):
                    raise ValueError("vae2_b200: HD_Z expects z as a list of 4 latent maps")
                zs = list(z)
            else:
                if not (torch.is_tensor(z) and z.dim() == 4 and z.shape[2:] == (1, 1)):
                    raise ValueError("vae2_b200: without HD_Z, z is one tensor [B, Z, 1, 1]")
                zs = [z]
            code = torch.randn(x.shape[0], self.z_dim, 1, 1, device=x.device).detach()   # reference :456
            inputs += zs + [code]
        x1, x2, x3 = self._run(inputs)
        return x1, x2, x3

    def sample_k(self, x, z, code):
        """K latent draws per context clip in ONE pass (reference: lib/core/function.py:124-146 runs the
        whole wrapper once per draw).  x: [B, 3L, H, W] context clips; z: 4 maps [K*B, Z, H_i, W_i]; code: [K*B, Z, 1, 1];
        draw k of clip b sits at row k*B + b.  The encoder trunk up to transition3 does not depend on z: it runs once per
        clip and is replicated; returns (x1, x2, x3) of shape [K*B, 3L, H, W].  Inference only (eval or no_grad)."""
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError("vae2_b200: sample_k is an inference path (call under torch.no_grad() or in eval mode)")
        KB, B = z[0].shape[0], x.shape[0]
        assert KB % B == 0 and code.shape[0] == KB and len(z) == 4
        with torch.no_grad():
            return self._run([x] + list(z) + [code], tag=("ksample", KB // B), batch=KB)


class HighResolutionNetEDz(HighResolutionNet):
    """Posterior network q(z | xt, x3t) with per-branch 1x1 z-heads (reference :984-1122)."""

    def __init__(self, config, **kwargs):
        extra = config.MODEL.EXTRA
        L = config.TRAIN.CLIP_LENGTH
        super().__init__(config, in_channels={"": 3 * L * 3 if extra.IS_BASELINE else 3 * L * 2},
                         heads=0, enable_random_code=False, **kwargs)
        self.last_layer = self._make_z_layer()

    def _conv1_repeat(self):
        L = self.clip_length
        return {"": L * 3 if self.extra.IS_BASELINE else L * 2}

    def _make_z_layer(self):
        if not self.hd_z:      # reference :1023-1041: global average pool -> 1x1(+bias) -> BN -> ReLU -> 1x1(+bias)
            return nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), _c(self.last_inp_channels, 512, 1, bias=True), _b(512),
                                 nn.ReLU(inplace=True), _c(512, 2 * self.z_dim, 1, bias=True))
        layers = []
        for c in self.last_stage_channels:       # reference :1000-1022
            layers.append(nn.Sequential(_c(c, self.z_dim * 2, 1)) if c != self.z_dim * 2 else None)
        return nn.ModuleList(layers)

    def _record(self, rec, shapes, needs, tag):
        _, Cx, H, W = shapes[0]
        if not self.hd_z:      # reference :1107-1116: up-sample + concat the four branches, then the pooled head
            _, cat = self._emit_trunk(rec, "", rec.input(Cx, H, W, needs[0]), head_cat=True)
            h = rec.conv_bn(rec.pool(cat), self.last_layer[1], self.last_layer[2], relu=True)
            o = rec.conv(h, self.last_layer[4])
            rec.output(o)
            return [(o.C, 1, 1)]
        ys, _ = self._emit_trunk(rec, "", rec.input(Cx, H, W, needs[0]), head_cat=False)
        outs = [rec.conv(y, self.last_layer[i][0]) for i, y in enumerate(ys)]
        for o in outs:
            rec.output(o)
        return [(o.C, o.H, o.W) for o in outs]

    def forward(self, x, *args, **kwargs):
        outs = self._run([x])
        return list(outs) if self.hd_z else outs[0]


class HighResolutionNetDsc(HighResolutionNet):
    """Sequence / frame discriminator with a 1-channel head (reference :1125-1183)."""

    def __init__(self, config, is_sequence, **kwargs):
        L = config.TRAIN.CLIP_LENGTH
        super().__init__(config, in_channels={"": 3 * L if is_sequence else 3}, heads=0,
                         enable_random_code=False, **kwargs)
        self.is_sequence = is_sequence
        self.last_layer = _head(self.last_inp_channels, 1, config.MODEL.EXTRA.FINAL_CONV_KERNEL)

    def _conv1_repeat(self):
        return {"": self.clip_length} if self.is_sequence else None

    def _record(self, rec, shapes, needs, tag):
        if isinstance(tag, tuple) and tag[0] == "groups":
            # stacked calls: group gi = channel window [coff, coff + Cin) of input #slot, samples gi*Bg .. (gi+1)*Bg
            _, Cin, groups = tag
            Bg, _, H, W = shapes[0]
            x = rec.plan.new_act(Cin, H, W, name="in_groups")
            for gi, (slot, coff) in enumerate(groups):
                rec.group_input(slot, Cin, H, W, shapes[slot][1], coff, x, gi, Bg, needs[slot])
            rec.n_in = len(shapes)
        else:
            _, Cx, H, W = shapes[0]
            x = rec.input(Cx, H, W, needs[0])
        _, cat = self._emit_trunk(rec, "", x)
        o = self._emit_head(rec, self.last_layer, cat)
        rec.output(o)
        return [(1, H, W)]

    def forward(self, x, *args, **kwargs):
        return self._run([x])[0]

    def forward_groups(self, sources):
        """Several calls of this discriminator as ONE stacked pass.  ``sources`` lists, in the
        reference's call order, (tensor [B, C, H, W], channel offset): call i sees channels [off, off + Cin) of its
        tensor, exactly what ``D(x[:, off:off + Cin])`` would.  Returns [len(sources) * B, 1, H, W]; rows
        [i*B, (i+1)*B) are call i's output.  Each call keeps its own BatchNorm batch statistics, running statistics
        are updated call by call and num_batches_tracked advances by len(sources) -- bit-for-bit the statistics
        grouping of the reference's sequential calls (lib/utils/utils.py:114-119, 259-267)."""
        cin = self.conv1.in_channels
        tensors, groups = [], []
        for t, off in sources:
            slot = next((i for i, u in enumerate(tensors) if u is t), None)
            if slot is None:
                tensors.append(t)
                slot = len(tensors) - 1
            assert t.shape == tensors[0].shape and off + cin <= t.shape[1]
            groups.append((slot, int(off)))
        B = tensors[0].shape[0]
        return self._run(tensors, tag=("groups", cin, tuple(groups)), batch=B * len(groups), stat_groups=len(groups))[0]


def get_encdec_model(cfg, **kwargs):
    model = HighResolutionNetED(cfg, **kwargs)
    model.init_weights(cfg.MODEL.PRETRAINED)
    return model


def get_D_sequence_model(cfg, **kwargs):
    model = HighResolutionNetDsc(config=cfg, is_sequence=True, **kwargs)
    model.init_weights(cfg.MODEL.PRETRAINED)
    return model


def get_D_frame_model(cfg, **kwargs):
    model = HighResolutionNetDsc(config=cfg, is_sequence=False, **kwargs)
    model.init_weights(cfg.MODEL.PRETRAINED)
    return model


def get_encz_model(cfg, **kwargs):
    model = HighResolutionNetEDz(cfg, **kwargs)
    model.init_weights(cfg.MODEL.PRETRAINED)
    return model
