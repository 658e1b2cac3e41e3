"""Same import surface as the reference's lib/models/__init__.py:11-13."""
from __future__ import absolute_import, division, print_function

# sub-modules this tree lacks resolve in a reference lib/ later on sys.path (models.seg_hrnet: the legacy
# segmentation net is outside the VAE^2 hot path, SURVEY.md §8, and stays the reference's own code)
__path__ = __import__("pkgutil").extend_path(__path__, __name__)

import models.enc_hrnet
import models.toy_fc
try:
    import models.seg_hrnet  # noqa: F401  (present only with a reference tree on sys.path)
except ImportError:
    pass
