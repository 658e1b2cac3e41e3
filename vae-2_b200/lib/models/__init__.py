"""Same import surface as the reference's lib/models/__init__.py:11-13."""
from __future__ import absolute_import, division, print_function

import models.seg_hrnet
import models.enc_hrnet
import models.toy_fc
