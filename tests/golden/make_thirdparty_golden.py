"""Re-derivation of the expectations in thirdparty_*.json (run here, no network): the published Brent optima with
scipy, the log1pExp right-hand sides in IEEE double, MurmurHash3 against sklearn's independent implementation."""
import json
import math
import os

from scipy.optimize import minimize_scalar
from sklearn.utils import murmurhash3_32

HERE = os.path.dirname(os.path.abspath(__file__))
FUNCS = {"sin": math.sin, "quintic": lambda x: (x - 1) * (x - 0.5) * x * (x + 0.5) * (x + 1),
         "math832": lambda x: 1e2 * math.sqrt(x) + 1e6 / x + 1e4 / math.sqrt(x)}

for c in json.load(open(os.path.join(HERE, "thirdparty_brent.json")))["cases"]:
    lo, hi = min(c["lo"], c["hi"]), max(c["lo"], c["hi"])
    if c["f"] == "math832":
        lo, hi = 1.0, 1e5
    x = minimize_scalar(FUNCS[c["f"]], bounds=(lo, hi), method="bounded", options={"xatol": 1e-13}).x
    print(f'{c["name"]:32s} published {c["expected"]!r:22} scipy {x!r:22} |diff| {abs(x - c["expected"]):.2e}')
for v in json.load(open(os.path.join(HERE, "thirdparty_murmur3.json")))["vectors"]:
    g = murmurhash3_32(v["data"].encode(), seed=v["seed"], positive=True)
    print(hex(v["seed"]), repr(v["data"][:16]), hex(g), "OK" if g == v["expected"] else "MISMATCH")
