from .default import _C as config
from .default import update_config, get_cfg_defaults, load_config
from .node import CfgNode

# sub-modules this tree lacks (the reference's control plane) resolve in a reference lib/ later on sys.path
__path__ = __import__("pkgutil").extend_path(__path__, __name__)
