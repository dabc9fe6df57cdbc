"""Extracts the /tf records of the reference's recording test/test_multiple_targets.bag (a data fixture, not source) with
the independent Python parser of tests/bagfile.py into tests/golden/bag_tf_records.npz.  Run in the build container, where
/root/reference exists:  python tests/golden/make_bag_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import bagfile  # noqa: E402

if __name__ == "__main__":
    rec = bagfile.parse_tf("/root/reference/test/test_multiple_targets.bag")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bag_tf_records.npz"), **rec)
    print("%d transforms in %d messages; frames: %s" % (len(rec["msg"]), int(rec["msg"].max()) + 1, sorted(set(rec["child_frame_id"]))))
    print("record time span: %.3f s" % ((int(rec["rec_sec"][-1]) - int(rec["rec_sec"][0])) + 1e-9 * (int(rec["rec_nsec"][-1]) - int(rec["rec_nsec"][0]))))
