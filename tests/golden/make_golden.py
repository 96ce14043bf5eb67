"""Regenerates tests/golden/golden_512.json from the CPU oracle on the reference's own sample inputs
(samples/images/512.png + samples/w_512.dat, copied here as 512.png / w_512.dat).

The reference records no expected outputs anywhere, so these known answers come from the
oracle; they pin the oracle against regressions and against the surveyor's independent numpy restatement
(a_NVF 2.852794, a_ME 34.90134 / 34.90300, corr_NVF 0.5856 / 0.5859, corr_ME 0.7379 / 0.7377).
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import oracle as o  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    img = util.load_512_gray(o)
    rgb = util.load_512_rgb()
    W = util.load_w512()
    g = {"inputs": {"gray_sha256": sha(img), "w_sha256": sha(W), "strength_psnr40": o.strength(40.0)}, "modes": {}}
    for name, opt in (("faithful", o.FAITHFUL), ("exact", o.EXACT), ("strict_f32", o.STRICT_F32)):
        m = {}
        pe = o.pred_error_mask(img, opt)
        m["Rx"] = pe["Rx"].tolist()
        m["rx"] = pe["rx"].tolist()
        m["coef"] = [float(c) for c in pe["coef"]]
        m["e_max_abs"] = float(np.abs(pe["e"]).max())
        m["e_crop"] = pe["e"][100:104, 200:204].astype(float).tolist()
        nv = o.nvf(img, opt)
        m["nvf_crop"] = nv[100:104, 200:204].astype(float).tolist()
        m["nvf_min"] = float(nv.min())
        for mask, mn in ((o.NVF, "nvf"), (o.ME, "me")):
            e = o.embed(img, W, 40.0, mask, o=opt)
            d = o.detect(e["out"], W, mask, o=opt)
            d0 = o.detect(img, W, mask, o=opt)
            e3 = o.embed(img, W, 40.0, mask, base=rgb, o=opt)
            q = e["out"].astype(np.uint8)
            dq = o.detect(q.astype(np.float32), W, mask, o=opt)
            m[mn] = {"a": e["a"], "corr_marked": d["corr"], "corr_clean": d0["corr"], "corr_marked_u8": dq["corr"],
                     "out_crop": e["out"][100:104, 200:204].astype(float).tolist(),
                     "out_u8_sha256": sha(q), "a_rgb": e3["a"],
                     "psnr": float(10 * np.log10(255.0 ** 2 / np.mean((e["out"].astype(np.float64) - img) ** 2)))}
        g["modes"][name] = m
    # u8 video frame path on the rounded luma
    y = np.rint(img).astype(np.uint8)
    st, out, a = o.embed_frame_u8(y, W, 40.0, o.ME)
    st2, corr = o.detect_frame_u8(out, W, o.ME)
    g["video_u8"] = {"a": a, "corr": corr, "out_sha256": sha(out), "in_sha256": sha(y)}
    with open(os.path.join(HERE, "golden_512.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden_512.json")


if __name__ == "__main__":
    main()
