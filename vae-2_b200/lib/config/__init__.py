from .default import _C as config
from .default import update_config, get_cfg_defaults, load_config
from .node import CfgNode
