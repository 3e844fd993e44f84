#!/bin/bash
timeout 900 python tools/r2_traj_noise.py 2>&1 | tail -4
