"""not gpu: the rosbag v2.0 /tf reader (host/bag_reader.cpp, target_bag_read_tf of include/target_manager_c.h) against an
independent pure-Python writer / parser (tests/bagfile.py), against the committed golden records of the reference's own
recording test/test_multiple_targets.bag, and -- in the build container -- against that recording itself."""
import os

import numpy as np
import pytest

from tests import bagfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "bag_tf_records.npz")
REF_BAG = "/root/reference/test/test_multiple_targets.bag"


def _same(got, rec):
    assert len(got) == len(rec["msg"])
    for k in ("rec_sec", "rec_nsec", "msg", "seq", "sec", "nsec"):
        assert np.array_equal(got[k], rec[k]), k
    assert [s.decode() for s in got["child_frame_id"]] == [str(s) for s in rec["child_frame_id"]]
    assert [s.decode() for s in got["frame_id"]] == [str(s) for s in rec["frame_id"]]
    assert np.array_equal(got["pose"].view(np.uint64), np.ascontiguousarray(rec["pose"]).view(np.uint64))   # bit-exact doubles


def test_reader_matches_python_writer(tmp_path):
    from target_estimation_b200.manager import read_bag_tf
    rng = np.random.default_rng(3)
    msgs, t = [], 1000 * 10 ** 9
    for m in range(300):
        t += int(rng.integers(1, 30_000_000))
        trs = []
        for j in range(int(rng.integers(0, 4))):   # empty messages too
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            child = ["target_%d" % rng.integers(0, 5), "camera_link", "target_filt_3"][int(rng.integers(0, 3)) if j else 0]
            trs.append((m, (t - 500_000) // 10 ** 9, (t - 500_000) % 10 ** 9, "cam", child, np.r_[rng.normal(size=3), q]))
        msgs.append(((t // 10 ** 9, t % 10 ** 9), trs))
    other = [((1000 + k, 5), b"\x05\x00\x00\x00hello") for k in range(5)]
    path = str(tmp_path / "synthetic.bag")
    bagfile.write_bag(path, msgs, other_topic_messages=other, chunk_messages=37)
    rec = bagfile.parse_tf(path)
    assert len(rec["msg"]) == sum(len(trs) for _, trs in msgs) > 100
    _same(read_bag_tf(path), rec)
    # the old message type name has the same wire format; another topic name selects nothing
    bagfile.write_bag(path, msgs[:10], msg_type="tf/tfMessage")
    assert len(read_bag_tf(path)) == sum(len(trs) for _, trs in msgs[:10])
    assert len(read_bag_tf(path, topic="/tf_static")) == 0


def test_reader_errors(tmp_path):
    from target_estimation_b200.manager import read_bag_tf
    q = [0, 0, 0, 1]
    msgs = [((1000, 0), [(0, 999, 5, "cam", "target_1", [1, 2, 3] + q)])]
    p = str(tmp_path / "a.bag")
    with pytest.raises(RuntimeError, match="cannot open"):
        read_bag_tf(str(tmp_path / "missing.bag"))
    open(p, "wb").write(b"#ROSBAG V1.2\n")
    with pytest.raises(RuntimeError, match="not a V2.0 bag"):
        read_bag_tf(p)
    bagfile.write_bag(p, msgs, compression=b"bz2")
    with pytest.raises(RuntimeError, match="compressed"):
        read_bag_tf(p)
    bagfile.write_bag(p, msgs)
    data = open(p, "rb").read()
    cut = data.index(b"target_1") + 20            # inside the first message body
    open(p, "wb").write(data[:cut])
    with pytest.raises(RuntimeError, match="truncated"):
        read_bag_tf(p)


def test_golden_records_round_trip(tmp_path):
    """the reference's recording, rebuilt from the committed records, reads back bit-identically"""
    from target_estimation_b200.manager import read_bag_tf
    rec = np.load(GOLDEN)
    assert len(rec["msg"]) == 572 and sorted(set(rec["child_frame_id"])) == ["target_0", "target_1", "target_2"]
    p = str(tmp_path / "rebuilt.bag")
    bagfile.write_bag(p, bagfile.messages_from_records(rec))
    _same(read_bag_tf(p), rec)


@pytest.mark.skipif(not os.path.exists(REF_BAG), reason="the reference checkout exists only in the build container")
def test_reference_recording_matches_golden():
    from target_estimation_b200.manager import read_bag_tf
    _same(read_bag_tf(REF_BAG), np.load(GOLDEN))
