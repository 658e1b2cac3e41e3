"""Minimal yacs-compatible config node.

The reference reads its configuration through ``yacs.config.CfgNode``
(reference lib/config/default.py:13, 121-127).  yacs is not a dependency of
this package; this stand-in supports exactly what the reference callers use:
attribute *and* item access (``cfg.MODEL.EXTRA.HD_Z`` and ``extra['STAGE1']``,
reference lib/models/enc_hrnet.py:262-310), ``merge_from_file`` (yaml),
``merge_from_list`` (``KEY VAL`` pairs from the CLI), ``defrost``/``freeze``
and ``new_allowed`` sub-trees (reference lib/config/default.py:38).
"""
import ast
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init=None, new_allowed=False):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        object.__setattr__(self, "_new_allowed", new_allowed)
        for k, v in (init or {}).items():
            self[k] = self._wrap(v, new_allowed)

    @staticmethod
    def _wrap(v, new_allowed=False):
        if isinstance(v, dict) and not isinstance(v, CfgNode):
            return CfgNode(v, new_allowed=new_allowed)
        return v

    # attribute access -----------------------------------------------------
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self._frozen:
            raise AttributeError("config is frozen; cannot set {}".format(name))
        self[name] = self._wrap(value, self._new_allowed)

    # yacs surface -----------------------------------------------------------
    def defrost(self):
        self._set_frozen(False)

    def freeze(self):
        self._set_frozen(True)

    def is_frozen(self):
        return self._frozen

    def _set_frozen(self, flag):
        object.__setattr__(self, "_frozen", flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode(new_allowed=self._new_allowed)
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        return out

    def merge_from_other_cfg(self, other, _path=""):
        for k, v in other.items():
            here = _path + k
            if k not in self:
                if not self._new_allowed:
                    raise KeyError("Non-existent config key: {}".format(here))
                self[k] = self._wrap(copy.deepcopy(v), True)
            elif isinstance(self[k], CfgNode) and isinstance(v, dict):
                self[k].merge_from_other_cfg(v, here + ".")
            else:
                self[k] = self._wrap(copy.deepcopy(v), self._new_allowed)

    def merge_from_file(self, path):
        with open(path, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self.merge_from_other_cfg(loaded)

    def merge_from_list(self, opts):
        opts = list(opts or [])
        if len(opts) % 2:
            raise ValueError("Override list has odd length: {}".format(opts))
        for key, raw in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError("Non-existent config key: {}".format(key))
                node = node[p]
            leaf = parts[-1]
            if leaf not in node and not node._new_allowed:
                raise KeyError("Non-existent config key: {}".format(key))
            node[leaf] = self._wrap(_decode(raw), node._new_allowed)


def _decode(raw):
    if not isinstance(raw, str):
        return raw
    try:
        return ast.literal_eval(raw)
    except (ValueError, SyntaxError):
        return raw
