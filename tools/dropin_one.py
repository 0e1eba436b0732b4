"""Runs one drop-in case (tests/test_gpu_dropin.py CASES[i]) outside pytest, e.g. under compute-sanitizer."""
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from oracle import refharness as R  # noqa: E402
from test_gpu_dropin import CASES, DROPIN_LIB, GROUPS  # noqa: E402

R.load()
lib = C.CDLL(DROPIN_LIB)
lib.dropin_run.restype = C.c_int
lib.dropin_run.argtypes = [C.POINTER(R.ref_case_spec), C.c_void_p, C.c_int]
for i in map(int, sys.argv[1:]):
    spec = R.make_spec(**CASES[i])
    out = np.full(len(GROUPS), -1.0)
    rc = lib.dropin_run(C.byref(spec), out.ctypes.data, out.size)
    print(i, CASES[i], "rc", rc, dict(zip(GROUPS, out.tolist())), flush=True)
