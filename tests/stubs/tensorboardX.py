"""Stub of tensorboardX for the drop-in tests: records add_scalar calls."""


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step=None):
        self.scalars.append((tag, float(value), step))

    def close(self):
        pass
