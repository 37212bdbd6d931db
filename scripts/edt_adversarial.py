"""Timings of the exact EDT on inputs that defeat the bounded scans (the cliff): a volume with a
single background voxel, a volume whose only background is one face plane, and the porous-media volume for
reference.  Results are checked against the CPU oracle at 256^3.
    python scripts/edt_adversarial.py [edge=512]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import porespy_b200 as psb
from oracle import cpu as oc                      # checker only
from porespy_b200 import _device as dev
from porespy_b200 import _lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = _lib.context(0)


def cases(edge):
    one = torch.ones((edge,) * 3, dtype=torch.uint8, device="cuda")
    one[edge // 2, edge // 3, edge // 5] = 0
    yield "single background voxel", one
    face = torch.ones((edge,) * 3, dtype=torch.uint8, device="cuda")
    face[0] = 0
    yield "background = the z = 0 face only", face
    corner = torch.ones((edge,) * 3, dtype=torch.uint8, device="cuda")
    corner[:, :, 0] = 0
    corner[:, 0, :] = 0
    yield "background = the x = 0 and y = 0 faces", corner
    yield "blobs(0.6, 2)", psb.generators.blobs([edge] * 3, porosity=0.6, blobiness=2, seed=0, rng="philox", as_numpy=False)


for name, im in cases(256):
    d2 = dev.edt_sq(ctx, im.reshape(-1), im.shape).cpu().numpy().view(np.uint32).reshape(im.shape)
    assert np.array_equal(d2, oc.edt_sq(im.cpu().numpy())), name
for name, im in cases(S):
    flat = im.reshape(-1)
    for _ in range(2):
        dev.edt_run(ctx, flat, im.shape, as_f32=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dev.edt_run(ctx, flat, im.shape, as_f32=True)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"edge": S, "input": name, "edt_ms": round(e0.elapsed_time(e1) / 3, 3), "parity_256": True}), flush=True)
