"""Opt-in NVTX ranges: host logic only, no GPU needed."""
import os
import subprocess
import sys

import pytest

from _util import ROOT
from fm_for_online_recommendation_b200 import tracing


class Recorder:
    def __init__(self):
        self.events = []

    def range_push(self, name):
        self.events.append(("push", name))

    def range_pop(self):
        self.events.append(("pop",))


def test_wrap_brackets_the_call_and_survives_exceptions():
    class Model:
        def update_embedding(self, x, scale=1):
            """doc"""
            return x * scale

        def fit(self, x):
            raise ValueError("Nan contained")

        def helper(self):
            return 0

    rec = Recorder()
    assert tracing.wrap(Model, nvtx=rec) == ["update_embedding", "fit"]
    assert tracing.wrap(Model, nvtx=rec) == []                      # idempotent
    m = Model()
    assert m.update_embedding(3, scale=2) == 6 and Model.update_embedding.__doc__ == "doc"
    with pytest.raises(ValueError, match="Nan contained"):
        m.fit(1)
    assert m.helper() == 0
    assert rec.events == [("push", "Model.update_embedding"), ("pop",), ("push", "Model.fit"), ("pop",)]


def test_off_by_default_and_on_with_the_environment_variable():
    code = ("import fm_for_online_recommendation_b200 as p; "
            "print(hasattr(p.FMAdam.update_embedding, '__wrapped__'), hasattr(p.FM_FTRL.online_learning, '__wrapped__'))")
    for flag, want in (("0", "False False"), ("1", "True True")):
        env = dict(os.environ, FMB_NVTX=flag, PYTHONPATH=ROOT)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert out.stdout.strip() == want
