"""The example training loop (examples/minimal_training.py): stand-in networks train through the fused loss."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_minimal_training_loop_reduces_the_loss():
    spec = importlib.util.spec_from_file_location("minimal_training", os.path.join(ROOT, "examples", "minimal_training.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    history = mod.main(["--steps", "30", "--height", "64", "--width", "96", "--batch", "2"])
    assert all(h == h and h > 0 for h in history)          # finite
    assert min(history[-5:]) < 0.9 * history[0], history    # the networks learn something
