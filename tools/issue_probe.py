#!/usr/bin/env python3
"""Runs the FP64 issue-model probe (nb_probe_fp64_mix) and prints DFMA throughput per mix."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodygo_b200 import capi

L = capi.load()
L.nb_probe_fp64_mix.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
names = {0: "8 DFMA", 1: "8 DFMA + 2 ALU", 2: "8 DFMA + 4 ALU", 3: "8 DFMA + 8 ALU", 4: "8 DFMA + 16 ALU",
         5: "8 DFMA + 1 MUFU.RSQ64H", 6: "8 DFMA + 4 ALU + 1 MUFU",
         7: "DFMA 3 distinct regs", 8: "DMUL 2 distinct regs", 9: "DADD 2 distinct regs"}
for kind, nm in names.items():
    tf = C.c_double(0)
    rc = L.nb_probe_fp64_mix(0, kind, 4096, C.byref(tf))
    print(f"{nm:28s} rc={rc} {tf.value:7.2f} TFLOP/s", flush=True)
