#!/usr/bin/env python
"""Writes tests/data/probes/sunsky.hdr: a small synthetic Radiance RGBE lat-long probe (128x64: sky
gradient, a bright sun disc, dark ground) for tests/data/envmini.tin, so that the HDR-probe part of
the path (ProbeSample / ProbePdf / ProbeEval, probe.h:105-236) has a scene small enough to commit and
to ship to the GPU box (the reference's probes/vankleef.hdr snapshot is 97 MB)."""
import math
import os
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 128, 64


def rgbe(r, g, b):
    m = max(r, g, b)
    if m < 1e-32:
        return (0, 0, 0, 0)
    e = math.frexp(m)[1]
    s = 256.0 / (2.0 ** e)
    return (int(r * s), int(g * s), int(b * s), e + 128)


def pixel(x, y):
    # y = 0 is the top row of the file (zenith side)
    theta = math.pi * (y + 0.5) / H
    phi = 2.0 * math.pi * (x + 0.5) / W
    up = math.cos(theta)
    if up < 0.0:
        return (0.05, 0.045, 0.04)
    t = math.sqrt(up)
    r, g, b = 0.9 * (1 - t) + 0.15 * t, 0.95 * (1 - t) + 0.3 * t, 1.0 * (1 - t) + 0.8 * t
    # sun: 60 degrees up, at phi = 1 rad
    sx, sy, sz = math.sin(math.pi / 6) * math.cos(1.0), math.cos(math.pi / 6), math.sin(math.pi / 6) * math.sin(1.0)
    dx, dy, dz = math.sin(theta) * math.cos(phi), up, math.sin(theta) * math.sin(phi)
    c = dx * sx + dy * sy + dz * sz
    if c > 0.985:
        r, g, b = r + 400.0, g + 360.0, b + 300.0
    elif c > 0.95:
        k = (c - 0.95) / 0.035
        r, g, b = r + 6.0 * k, g + 5.0 * k, b + 4.0 * k
    return (r, g, b)


def main():
    out = os.path.join(ROOT, "tests", "data", "probes", "sunsky.hdr")
    with open(out, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (H, W))
        for y in range(H):
            row = [rgbe(*pixel(x, y)) for x in range(W)]
            f.write(struct.pack("BBBB", 2, 2, W >> 8, W & 255))
            for c in range(4):
                # literal (non-run) packets of up to 128 bytes
                data = bytes(p[c] for p in row)
                for i in range(0, W, 128):
                    chunk = data[i:i + 128]
                    f.write(bytes([len(chunk)]) + chunk)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
