import sys, importlib, torch
sys.path.insert(0, '.')
import b200asr
from tests.helpers import rel_err
from tests.test_gpu_parity import _attention_case
ops = importlib.import_module(b200asr.__name__ + ".ops")
ops.config.set(attn="tf32", attn_bwd="fp32")
for case in [(2, 8, 100, 200, 64, 64, "keypad"), (1, 1, 200, 200, 64, 64, "keypad"), (1, 2, 300, 448, 64, 64, "keypad"), (2, 2, 257, 400, 64, 64, "none"), (3, 4, 13, 13, 32, 32, "causal+keypad")]:
    pairs = _attention_case(ops, *case)
    print(case, " ".join("%.2e" % rel_err(a, b) for a, b in pairs))
