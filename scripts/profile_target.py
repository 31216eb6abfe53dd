"""Small, fixed workload for ncu captures (never a bench number): python scripts/profile_target.py [pair|planar|nn]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icpslam_b200 import synth, registration as R

mode = sys.argv[1] if len(sys.argv) > 1 else "pair"
if mode == "planar":
    _, _, sc = synth.planar_stream(3, 2)
    src, tgt, preset = sc[1], sc[0], R.PRESET_ODOMETER
else:
    _, _, sw = synth.sweep_sequence(4, 2)
    src, tgt, preset = sw[1], sw[0], R.PRESET_MAPPER
reg = R.Registration(preset=preset)
reg.setInputTarget(tgt)
reg.setInputSource(src)
if mode == "nn":
    for _ in range(3):
        reg.nearestKSearch1(src)
else:
    for _ in range(2):
        reg.align()
    print(reg.iterations, reg.getFitnessScore())
