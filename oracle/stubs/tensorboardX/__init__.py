"""Harness-side stand-in for ``tensorboardX`` (absent from this image).  TEST INFRASTRUCTURE ONLY.
util/initPara.py:217 builds ``SummaryWriter(log_dir=...)``; main.py:43 closes it; the train loops call add_scalar."""


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step=None, *a, **k):
        self.scalars.append((tag, float(value), step))

    def close(self):
        pass

    def __getattr__(self, name):          # add_histogram, add_text, flush ... : accepted and ignored
        return lambda *a, **k: None
