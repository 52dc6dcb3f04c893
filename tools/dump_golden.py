"""For anyone WITH opencv-contrib-python: dump real cv2.optflow DualTVL1 flows for the committed golden inputs so
that the restated oracle (oracle/tvl1_oracle.c) can be pinned against genuine OpenCV output.

    python tools/dump_golden.py            # writes tests/golden/tvl1_pairs_cv2optflow.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def main():
    import cv2
    if not hasattr(cv2, "optflow"):
        sys.exit("this cv2 build has no optflow module (install opencv-contrib-python)")
    g = np.load(ROOT / "tests" / "golden" / "tvl1_pairs.npz")
    out = {}
    for key in g.files:
        if key.endswith("__I0"):
            name = key[:-4]
            m = cv2.optflow.createOptFlow_DualTVL1()
            import ast
            params = ast.literal_eval(str(g[f"{name}__params"]))
            setters = dict(lambda_="setLambda", tau="setTau", theta="setTheta", nscales="setScalesNumber",
                           warps="setWarpingsNumber", epsilon="setEpsilon", inner="setInnerIterations",
                           outer="setOuterIterations")
            for k, v in params.items():
                getattr(m, setters[k])(v)
            out[f"{name}__flow_cv2"] = m.calc(g[f"{name}__I0"], g[f"{name}__I1"], None)
            d = out[f"{name}__flow_cv2"] - g[f"{name}__flow_em0"]
            print(name, "mean EPE vs restated oracle:", float(np.sqrt((d ** 2).sum(-1)).mean()))
    np.savez_compressed(ROOT / "tests" / "golden" / "tvl1_pairs_cv2optflow.npz", **out)


if __name__ == "__main__":
    main()
