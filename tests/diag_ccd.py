"""Helper (not a test): how many bodies the CCD criterion queues on the headline scene."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld
w = PhysicsWorld(scenes.many_pyramids_label())
for n in (1, 5, 50):
    w.step(n)
    st = w.debug_read("state", np.int32)
    print("steps", n, "nccd", st[38], "ccd_total", st[39])
