def __getattr__(name):
    def _nop(*a, **k):
        return None
    return _nop
