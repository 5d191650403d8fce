#!/usr/bin/env python3
"""Converts images (JPEG / PNG / anything PIL reads) to the binary PPM the `cnn` CLI of this
repository loads (host/Context.cpp: the reference decodes with its vendored stb_image, the image
codecs are outside the hot path), and PPM results back to PNG.

    python tools/img2ppm.py in.jpg out.ppm           # one file, either direction by extension
    python tools/img2ppm.py --dir samples/ [--out samples_ppm/]
        converts every *_large.jpg / *_small.jpg (the reference's training-sample naming,
        src/Main_cl.cpp:267-301) to *_large.ppm / *_small.ppm
"""
import argparse
import os
import sys

from PIL import Image


def convert(src, dst):
    img = Image.open(src).convert("RGB")
    if dst.lower().endswith((".ppm", ".pnm")):
        img.save(dst, format="PPM")
    else:
        img.save(dst)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src", nargs="?")
    ap.add_argument("dst", nargs="?")
    ap.add_argument("--dir")
    ap.add_argument("--out")
    a = ap.parse_args()
    if a.dir:
        out = a.out or a.dir
        os.makedirs(out, exist_ok=True)
        n = 0
        for f in sorted(os.listdir(a.dir)):
            stem, ext = os.path.splitext(f)
            if ext.lower() in (".jpg", ".jpeg", ".png") and stem.endswith(("_large", "_small")):
                convert(os.path.join(a.dir, f), os.path.join(out, stem + ".ppm"))
                n += 1
        print("converted %d sample images into %s" % (n, out))
        return 0
    if not a.src or not a.dst:
        ap.error("give SRC and DST, or --dir")
    convert(a.src, a.dst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
