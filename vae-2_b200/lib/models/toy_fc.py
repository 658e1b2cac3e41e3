"""Toy VAE^2 MLPs (reference lib/models/toy_fc.py:16-176), BASELINE config #1.

Three-layer perceptrons (10 -> 128 -> 128 -> 10, Z = 8) used by tools/toy_example.py to check
the ELBO plumbing.  There is no convolution or BN here: the linear layers are plain
torch.nn (plumbing), while the ELBO terms they feed run through the fused CUDA kernel
(core/criterion.py).  Same class names, attributes and state_dict keys as the reference.
"""
import logging

import torch
import torch.nn as nn

HID_DIM, Z_DIM, INPUT_DIM = 128, 8, 10
logger = logging.getLogger(__name__)


def _fc(i, o):
    return nn.Sequential(nn.Linear(i, o), nn.ReLU(inplace=True))


class toy_fc(nn.Module):
    def __init__(self, config):
        super().__init__()
        ex = config.MODEL.EXTRA
        self.is_baseline, self.baseline_mode = ex.IS_BASELINE, ex.BASELINE_MODE
        self.I_e_dim = INPUT_DIM * 2 if self.is_baseline else INPUT_DIM
        self.I_s_dim = self.v_dim = INPUT_DIM
        self.z_dim = 0 if self.baseline_mode == "DETERMINISTIC" else Z_DIM
        self.h1, self.h2 = _fc(self.I_e_dim, HID_DIM), _fc(HID_DIM, HID_DIM)
        self.output = nn.Linear(HID_DIM, self.v_dim)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.001)

    def _gen_code_map(self, feature, code=None):
        if code is not None:
            return code
        return torch.randn(feature.shape[0], self.z_dim, device=feature.device).detach()

    def forward(self, x, *args, **kwargs):
        return self.output(self.h2(self.h1(x)))


class toy_fc_EDz(toy_fc):
    def __init__(self, config):
        super().__init__(config)
        self.h1 = _fc(self.I_e_dim + self.v_dim, HID_DIM)
        self.output = nn.Linear(HID_DIM, Z_DIM * 2)


class toy_fc_ED(toy_fc):
    def __init__(self, config):
        super().__init__(config)
        zin = self.z_dim if self.is_baseline else 2 * self.z_dim
        self.h1 = _fc(self.I_e_dim + zin, HID_DIM)
        for p, odim in (("decp_", self.I_e_dim), ("decf_", self.v_dim)):
            setattr(self, p + "h1", _fc(self.I_s_dim + self.z_dim, HID_DIM))
            setattr(self, p + "h2", _fc(HID_DIM, HID_DIM))
            setattr(self, p + "output", nn.Linear(HID_DIM, odim))

    def _net(self, p, x):
        return getattr(self, p + "output")(getattr(self, p + "h2")(getattr(self, p + "h1")(x)))

    def _encoder_forward(self, x, z):
        cz = self._gen_code_map(x, z)
        cr = self._gen_code_map(x)          # drawn after z's code, as the reference does (:109-110)
        if self.is_baseline:
            return self._net("", x if self.baseline_mode == "DETERMINISTIC" else torch.cat([x, cz], -1))
        return self._net("", torch.cat([x, cz, cr], -1))

    def _decoder(self, p, x, z):
        cz = self._gen_code_map(x, z)
        if self.is_baseline and self.baseline_mode == "DETERMINISTIC":
            return self._net(p, x)
        return self._net(p, torch.cat([x, cz], -1))

    def forward(self, x, z=None, *args, **kwargs):
        x2 = self._encoder_forward(x, z)
        x1 = self._decoder("decp_", x2, z)
        x3 = self._decoder("decf_", x2, z)
        return x1, x2, x3


class toy_fc_Dsc(toy_fc):
    def __init__(self, config):
        super().__init__(config)
        self.h1 = _fc(self.I_s_dim, HID_DIM)
        self.output = nn.Linear(HID_DIM, 1)


def get_encdec_model(config):
    m = toy_fc_ED(config)
    m.init_weights()
    return m


def get_encz_model(config):
    m = toy_fc_EDz(config)
    m.init_weights()
    return m


def get_D_model(config):
    m = toy_fc_Dsc(config)
    m.init_weights()
    return m
