import os
import sys

_LIB = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _LIB not in sys.path:
    sys.path.insert(0, _LIB)
from _engine_loader import engine  # noqa: E402

EngineModule = engine().EngineModule
