"""Test infrastructure: an independent pure-Python rosbag v2.0 writer and /tf parser (struct only, no ROS), used to check
the C++ reader (target_estimation_b200/host/bag_reader.cpp) and to rebuild a bag from the committed golden records of the
reference's recording (tests/golden/bag_tf_records.npz) on machines where /root/reference does not exist."""
import struct

import numpy as np

MAGIC = b"#ROSBAG V2.0\n"
TF_DEF = b"geometry_msgs/TransformStamped[] transforms\n"


def _fields(d):
    out = b""
    for k, v in d.items():
        kv = k.encode() + b"=" + v
        out += struct.pack("<I", len(kv)) + kv
    return out


def _record(header, data):
    h = _fields(header)
    return struct.pack("<I", len(h)) + h + struct.pack("<I", len(data)) + data


def _string(s):
    b = s.encode() if isinstance(s, str) else bytes(s)
    return struct.pack("<I", len(b)) + b


def tf_message(transforms):
    """transforms: list of (seq, sec, nsec, frame_id, child_frame_id, pose7)"""
    body = struct.pack("<I", len(transforms))
    for seq, sec, nsec, frame, child, pose in transforms:
        body += struct.pack("<III", int(seq), int(sec), int(nsec)) + _string(frame) + _string(child) + struct.pack("<7d", *[float(x) for x in pose])
    return body


def write_bag(path, messages, topic="/tf", msg_type="tf2_msgs/TFMessage", compression=b"none", other_topic_messages=(), chunk_messages=64):
    """messages: list of ((rec_sec, rec_nsec), [transform, ...]); other_topic_messages: raw (rec_time, bytes) on a second
    connection (must be skipped by a /tf reader).  Several chunks of `chunk_messages` messages each."""
    conn_hdr = {"topic": topic.encode(), "type": msg_type.encode(), "md5sum": b"94810edda583a504dfda3829e70d7eec", "message_definition": TF_DEF}
    conn0 = _record({"op": b"\x07", "conn": struct.pack("<I", 0), "topic": topic.encode()}, _fields(conn_hdr))
    conn1 = _record({"op": b"\x07", "conn": struct.pack("<I", 1), "topic": b"/other"},
                    _fields({"topic": b"/other", "type": b"std_msgs/String", "md5sum": b"992ce8a1687cec8c8bd883ec73ca41d1", "message_definition": b"string data\n"}))
    items = [(t, 0, tf_message(trs)) for t, trs in messages] + [(t, 1, raw) for t, raw in other_topic_messages]
    items.sort(key=lambda it: it[0][0] * 10 ** 9 + it[0][1])
    chunks = []
    for c0 in range(0, max(len(items), 1), chunk_messages):
        data = conn0 + conn1 if c0 == 0 else b""
        for (sec, nsec), conn, body in items[c0:c0 + chunk_messages]:
            data += _record({"op": b"\x02", "conn": struct.pack("<I", conn), "time": struct.pack("<II", int(sec), int(nsec))}, body)
        chunks.append(_record({"op": b"\x05", "compression": compression, "size": struct.pack("<I", len(data))}, data))
        # an index-data record after each chunk, as rosbag writes them (content irrelevant to a sequential reader)
        chunks.append(_record({"op": b"\x04", "ver": struct.pack("<I", 1), "conn": struct.pack("<I", 0), "count": struct.pack("<I", 0)}, b""))
    body = b"".join(chunks)
    hdr_fields = _fields({"op": b"\x03", "index_pos": struct.pack("<Q", len(MAGIC) + 4096 + len(body)), "conn_count": struct.pack("<I", 2),
                          "chunk_count": struct.pack("<I", len(chunks) // 2)})
    pad = 4096 - 4 - len(hdr_fields) - 4
    header = struct.pack("<I", len(hdr_fields)) + hdr_fields + struct.pack("<I", pad) + b" " * pad
    with open(path, "wb") as f:
        f.write(MAGIC + header + body + conn0 + conn1)


def _parse_fields(b):
    d, p = {}, 0
    while p < len(b):
        (l,) = struct.unpack_from("<I", b, p); p += 4
        k, v = b[p:p + l].split(b"=", 1); p += l
        d[k.decode()] = v
    return d


def parse_tf(path, topic="/tf"):
    """independent parser -> dict of arrays (one row per transform)"""
    data = open(path, "rb").read()
    assert data.startswith(MAGIC)
    rows, wanted, n_msg = [], {}, [0]

    def walk(buf):
        p = 0
        while p < len(buf):
            (hl,) = struct.unpack_from("<I", buf, p); p += 4
            h = _parse_fields(buf[p:p + hl]); p += hl
            (dl,) = struct.unpack_from("<I", buf, p); p += 4
            body = buf[p:p + dl]; p += dl
            op = h["op"][0]
            if op == 5:
                assert h["compression"] == b"none"
                walk(body)
            elif op == 7:
                ch = _parse_fields(body)
                wanted[struct.unpack("<I", h["conn"])[0]] = h["topic"].decode() == topic and ch.get("type") in (b"tf2_msgs/TFMessage", b"tf/tfMessage")
            elif op == 2 and wanted.get(struct.unpack("<I", h["conn"])[0]):
                rs, rn = struct.unpack("<II", h["time"])
                (cnt,) = struct.unpack_from("<I", body, 0); q = 4
                for _ in range(cnt):
                    seq, sec, nsec = struct.unpack_from("<III", body, q); q += 12
                    (l,) = struct.unpack_from("<I", body, q); q += 4; frame = body[q:q + l].decode(); q += l
                    (l,) = struct.unpack_from("<I", body, q); q += 4; child = body[q:q + l].decode(); q += l
                    pose = struct.unpack_from("<7d", body, q); q += 56
                    rows.append((rs, rn, n_msg[0], seq, sec, nsec, frame, child, pose))
                n_msg[0] += 1

    walk(data[len(MAGIC):])
    return {"rec_sec": np.array([r[0] for r in rows], dtype=np.uint32), "rec_nsec": np.array([r[1] for r in rows], dtype=np.uint32),
            "msg": np.array([r[2] for r in rows], dtype=np.uint32), "seq": np.array([r[3] for r in rows], dtype=np.uint32),
            "sec": np.array([r[4] for r in rows], dtype=np.uint32), "nsec": np.array([r[5] for r in rows], dtype=np.uint32),
            "frame_id": np.array([r[6] for r in rows]), "child_frame_id": np.array([r[7] for r in rows]),
            "pose": np.array([r[8] for r in rows], dtype=np.float64).reshape(-1, 7)}


def messages_from_records(rec):
    """records dict (parse_tf / golden npz) -> write_bag's message list"""
    msgs = []
    for i in range(len(rec["msg"])):
        tr = (int(rec["seq"][i]), int(rec["sec"][i]), int(rec["nsec"][i]), str(rec["frame_id"][i]), str(rec["child_frame_id"][i]), rec["pose"][i])
        if msgs and msgs[-1][2] == int(rec["msg"][i]):
            msgs[-1][1].append(tr)
        else:
            msgs.append(((int(rec["rec_sec"][i]), int(rec["rec_nsec"][i])), [tr], int(rec["msg"][i])))
    return [(t, trs) for t, trs, _ in msgs]
