"""Golden vectors for the input preparation, produced by the REFERENCE's own ``data_util.pad_volume``
(data_util.py:814-828).  ``data_util`` cannot be imported here (it pulls in monai, SimpleITK, nibabel and five modules the
reference does not ship), so the function's source is cut out of the file with ``ast`` and executed as is.  The SimpleITK
resample (VolumeDataset.py:236-259) cannot be run in this container; its restatement stays unpinned (oracle/prepare.py).

    python -m tests.golden.make_prepare_golden        # build container only; writes tests/golden/prepare_golden.npz
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("COMA_REFERENCE", "/root/reference")

# (input shape [1, z, y, x], target_size) -- odd paddings, an over-long y axis (cropped), over-long x/z (kept), exact fit
CASES = {
    "odd":    ((1, 5, 6, 7), (8, 8, 8)),
    "crop_y": ((1, 6, 11, 5), (8, 8, 8)),
    "long_xz": ((1, 10, 4, 9), (8, 6, 8)),
    "ragged": ((1, 3, 9, 4), (7, 5, 6)),
    "fit":    ((1, 4, 8, 8), (8, 8, 8)),
}


def case_input(name):
    shape, _ = CASES[name]
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    return torch.randn(shape, generator=g)


def reference_pad_volume():
    src = open(os.path.join(REFERENCE, "data_util.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "pad_volume")
    scope = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "data_util.py", "exec"), scope)
    return scope["pad_volume"]


def main():
    pad_volume = reference_pad_volume()
    out = {}
    for name, (_, target) in CASES.items():
        out[name] = pad_volume(target)(case_input(name)).numpy()
    np.savez_compressed(os.path.join(HERE, "prepare_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
