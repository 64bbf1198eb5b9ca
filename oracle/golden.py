"""TEST INFRASTRUCTURE — loader for the committed golden fixtures (tests/golden/*.pt)."""
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name), map_location="cpu", weights_only=False)
