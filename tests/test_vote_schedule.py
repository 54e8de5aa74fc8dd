"""The parallel schedule of the iterative region voting (adcensus_b200/csrc/k_vote.cu) restated on the CPU:
tools/sim/push_sim.c keeps one histogram per pending pixel and runs derive / push rounds to the fixed point of every sweep;
it must end in exactly the post-voting map of the sequential reference (oracle) on the Cone pair.  This pins the
*algorithm* the GPU kernel implements without a GPU; the kernel itself is checked by tests/test_gpu_parity.py."""
import os
import subprocess
from pathlib import Path

import numpy as np

import adc_testlib as T

SIM = T.REPO / "tools" / "sim"


def test_push_schedule_equals_sequential_voting(tmp_path):
    left, right = T.load_cone()
    o = T.Oracle(450, 375)
    o.begin(left, right)
    o.run_to("OUTLIER")
    o.tap("DISP_L").tofile(tmp_path / "disp.bin")
    o.tap("ARMS").tofile(tmp_path / "arms.bin")
    o.tap("SUPCNT_H").tofile(tmp_path / "suph.bin")
    o.tap("MISMATCHES").astype(np.int32).tofile(tmp_path / "mm.bin")
    o.tap("OCCLUSIONS").astype(np.int32).tofile(tmp_path / "oc.bin")
    o.run_to("VOTE")
    o.tap("DISP_L").tofile(tmp_path / "disp_vote.bin")
    exe = tmp_path / "push_sim"
    subprocess.run(["gcc", "-O2", "-w", "-o", str(exe), str(SIM / "push_sim.c"), "-lm"], check=True)
    r = subprocess.run([str(exe)], env=dict(os.environ, ADC_SIM_DIR=str(tmp_path)), capture_output=True, text=True, check=True)
    assert "mismatch_vs_ref 0" in r.stdout, r.stdout
    # the work the schedule does (the numbers DESIGN.md quotes): one region scan per pending pixel, then pushes
    fields = dict(zip(r.stdout.split()[1::2], r.stdout.split()[2::2]))
    assert int(fields["rounds"]) < 200 and int(fields["derives"]) < 100000, r.stdout
