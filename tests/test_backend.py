"""WebSocket JSON backend: wire format of src/app/svo_slam_backend.cpp:18-110 and a real RFC 6455
round trip over localhost, against a stand-in slam object (the payload builders only use the StereoSlam getters)."""
import base64
import hashlib
import json
import os
import socket
import struct
import threading

import numpy as np
import pytest

from stereo_svo_slam_b200 import backend, capi, synth


class FakeKps:
    def __init__(self, n):
        self.kps3d = np.arange(3 * n, dtype=np.float32).reshape(n, 3) * 0.5
        self.info = np.zeros(n, capi.KPINFO_DTYPE)
        self.info["color"] = np.arange(3 * n, dtype=np.uint8).reshape(n, 3)


class FakeFrame:
    def __init__(self, pose, n):
        self.pose, self.kps = np.array(pose, np.float32), FakeKps(n)


class FakeSlam:
    def __init__(self):
        self.kfs = [FakeFrame([0, 0, 0, 0, 0, 0], 2), FakeFrame([1, 2, 3, 0.1, -0.2, 0.3], 3)]
        self.frame = FakeFrame([1.5, 2.5, 3.5, 0.05, 0.02, -0.01], 1)

    def get_keyframes(self):
        return self.kfs

    def get_frame(self):
        return self.frame

    def get_trajectory(self):
        return np.array([[0, 0, 0, 0, 0, 0], [1, 2, 3, 0.1, -0.2, 0.3]], np.float32)


def test_robot_angles_matches_reference_order():
    # PoseManager::get_robot_angles (pose_manager.cpp:45-59): Rodrigues(Rz * (Rx * Ry)); single-axis rotations are unchanged
    assert np.allclose(backend.robot_angles([0, 0, 0, 0.3, 0, 0]), [0.3, 0, 0])
    assert np.allclose(backend.robot_angles([0, 0, 0, 0, 0, -0.2]), [0, 0, -0.2])
    a = backend.robot_angles([0, 0, 0, 0.1, -0.2, 0.3])
    R = synth._rodrigues(a)
    want = synth._rodrigues([0, 0, 0.3]) @ (synth._rodrigues([0.1, 0, 0]) @ synth._rodrigues([0, -0.2, 0]))
    assert np.allclose(R, want, atol=1e-9)


def test_payloads_follow_the_reference_format():
    s = FakeSlam()
    kfs = json.loads(backend.answer(s, "/keyframes", "get"))
    assert backend.answer(s, "/keyframes", "something else") is None              # svo_slam_backend.cpp:27
    assert len(kfs) == 2 and set(kfs[0]) == {"pose", "keypoints", "colors"}
    assert set(kfs[1]["pose"]) == {"x", "y", "z", "rx", "ry", "rz"} and kfs[1]["pose"]["x"] == 1.0
    assert kfs[1]["keypoints"][2] == {"x": 3.0, "y": 3.5, "z": 4.0} and kfs[1]["colors"][1] == {"r": 3, "g": 4, "b": 5}
    p = json.loads(backend.answer(s, "ws://localhost:8001/pose", "anything"))
    assert set(p) == {"pose"} and p["pose"]["z"] == 3.5
    t = json.loads(backend.answer(s, "/trajectory", ""))
    assert t["trajectory"][6:9] == [1.0, 2.0, 3.0] and len(t["trajectory"]) == 12   # raw pose vectors, flat (:88-97)
    assert backend.answer(s, "/unknown", "get") is None
    assert " " not in backend.answer(s, "/pose", "")                               # QJsonDocument::Compact


def _ws_client(port, resource, message):
    c = socket.create_connection(("127.0.0.1", port), timeout=5)
    key = base64.b64encode(os.urandom(16))
    c.sendall(b"GET " + resource.encode() + b" HTTP/1.1\r\nHost: localhost\r\nUpgrade: websocket\r\nConnection: Upgrade\r\n"
              b"Sec-WebSocket-Key: " + key + b"\r\nSec-WebSocket-Version: 13\r\n\r\n")
    head = b""
    while b"\r\n\r\n" not in head:
        head += c.recv(1)
    assert b"101" in head.split(b"\r\n")[0]
    want = base64.b64encode(hashlib.sha1(key + b"258EAFA5-E914-47DA-95CA-C5AB0DC85B11").digest())
    assert want in head
    payload, mask = message.encode(), os.urandom(4)
    c.sendall(bytes([0x81, 0x80 | len(payload)]) + mask + bytes(b ^ mask[i % 4] for i, b in enumerate(payload)))
    b0, b1 = c.recv(2)
    n = b1 & 127
    if n == 126:
        n = struct.unpack(">H", c.recv(2))[0]
    elif n == 127:
        n = struct.unpack(">Q", c.recv(8))[0]
    data = b""
    while len(data) < n:
        data += c.recv(n - len(data))
    c.sendall(bytes([0x88, 0x80]) + os.urandom(4))
    c.close()
    assert b0 == 0x81
    return data.decode()


@pytest.mark.parametrize("resource,message,key", [("/keyframes", "get", None), ("/pose", "x", "pose"), ("/trajectory", "x", "trajectory")])
def test_websocket_round_trip(resource, message, key):
    srv = backend.WebSocketServer(FakeSlam(), port=0, host="127.0.0.1")
    got = {}
    t = threading.Thread(target=lambda: got.setdefault("text", _ws_client(srv.port, resource, message)))
    t.start()
    for _ in range(200):                      # the tracking thread polls between frames
        srv.serve_pending(timeout=0.05)
        if not t.is_alive():
            break
    t.join(timeout=5)
    srv.close()
    msg = json.loads(got["text"])
    assert (isinstance(msg, list) and len(msg) == 2) if key is None else key in msg


def _handshake(port, resource="/keyframes"):
    c = socket.create_connection(("127.0.0.1", port), timeout=5)
    key = base64.b64encode(os.urandom(16))
    c.sendall(b"GET " + resource.encode() + b" HTTP/1.1\r\nHost: localhost\r\nUpgrade: websocket\r\nConnection: Upgrade\r\n"
              b"Sec-WebSocket-Key: " + key + b"\r\nSec-WebSocket-Version: 13\r\n\r\n")
    return c


def test_websocket_server_survives_bad_clients():
    """The server runs on the tracking thread: a viewer that disconnects in the middle of an answer, announces a huge frame
    or never reads must cost its own connection, never an exception in (or a stall of) the frame loop."""
    class BigSlam(FakeSlam):
        pass
    slam = BigSlam()
    srv = backend.WebSocketServer(slam, port=0)           # loopback by default
    assert srv.srv.getsockname()[0] == "127.0.0.1"
    # 1. request, then vanish without reading the answer
    c = _handshake(srv.port)
    for _ in range(20):
        srv.serve_pending(timeout=0.02)
    mask = os.urandom(4)
    c.sendall(bytes([0x81, 0x83]) + mask + bytes(b ^ mask[i % 4] for i, b in enumerate(b"get")))
    c.setsockopt(socket.SOL_SOCKET, socket.SO_LINGER, struct.pack("ii", 1, 0))   # RST on close
    c.close()
    for _ in range(20):
        srv.serve_pending(timeout=0.02)                   # must not raise BrokenPipeError / ConnectionResetError
    assert srv.clients == []
    # 2. a frame header that announces 1 GiB: dropped, nothing buffered
    c = _handshake(srv.port)
    for _ in range(20):
        srv.serve_pending(timeout=0.02)
    c.sendall(bytes([0x81, 0x80 | 127]) + struct.pack(">Q", 1 << 30) + os.urandom(4) + b"x" * 100)
    for _ in range(20):
        srv.serve_pending(timeout=0.02)
    assert srv.clients == []
    c.close()
    # 3. the server still answers a well-behaved client afterwards
    got = {}
    t = threading.Thread(target=lambda: got.setdefault("text", _ws_client(srv.port, "/pose", "x")))
    t.start()
    for _ in range(200):
        srv.serve_pending(timeout=0.05)
        if not t.is_alive():
            break
    t.join(timeout=5)
    srv.close()
    assert "pose" in json.loads(got["text"])


def test_slam_accelerator_views_follow_the_cython_wrapper():
    """slam_accelerator drop-in: attribute-style access of src/python/wrapper/slam_accelerator.pyx
    (draw_kps.py reads kps2d[i].x, info[i].color['r'], info[i].type == KeyPointType.KP_FAST, pose.x, id)."""
    from stereo_svo_slam_b200 import slam_accelerator as sa
    cs = sa.CameraSettings()
    cs.fx, cs.grid_width, cs.max_pyramid_levels = 470.0, 40, 5            # fields assigned one by one (src/python/main.py:28-46)
    assert cs.fx == 470.0 and cs.grid_width == 40 and cs.baseline == 0.0
    s = FakeSlam()
    f = s.kfs[1]
    f.id = 7
    f.kps.kps2d = np.array([[1.5, 2.5], [3.0, 4.0], [5.0, 6.0]], np.float32)
    f.kps.info["type"][1] = 1
    f.kps.info["keyframe_id"][2] = 3
    f.image = lambda kind, level: np.full((4 >> level, 6 >> level), 9, np.uint8)
    v = sa.KeyFrame(f, levels=2)
    assert v.id == 7 and v.pose.x == 1.0 and abs(v.pose.rz - 0.3) < 1e-6
    assert len(v.kps.kps2d) == 3 and v.kps.kps2d[0].x == 1.5 and v.kps.kps3d[2].z == 4.0
    assert v.kps.info[1].type == sa.KeyPointType.KP_EDGELET and v.kps.info[0].type == sa.KeyPointType.KP_FAST
    assert v.kps.info[1].color == {"r": 3, "g": 4, "b": 5} and v.kps.info[2].keyframe_id == 3
    assert [im.shape for im in v.stereo_image.left] == [(4, 6), (2, 3)] and len(v.stereo_image.right) == 1
    assert sa.StereoSlam(cs).get_frame() is None and sa.StereoSlam(cs).get_keyframes() == []
