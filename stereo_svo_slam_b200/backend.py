"""WebSocket JSON backend of the reference app, so that `src/qt-viewer` works against this library
unchanged.

Wire format = `SvoSlamBackend::text_message_received` (src/app/svo_slam_backend.cpp:18-110); the resource name of the
socket selects the answer, any text message triggers it (`keyframes` answers only to the message "get"):

    ws://host:8001/keyframes  ->  [{"pose": {x,y,z,rx,ry,rz}, "keypoints": [{x,y,z}, ...], "colors": [{r,g,b}, ...]}, ...]
    ws://host:8001/pose       ->  {"pose": {x,y,z,rx,ry,rz}}
    ws://host:8001/trajectory ->  {"trajectory": [x,y,z,rx,ry,rz, x,y,z,...]}

Angles of `pose` objects are `PoseManager::get_robot_angles` (src/lib/pose_manager.cpp:45-59: Rodrigues(Rz*(Rx*Ry))), the
trajectory carries the raw pose vectors (svo_slam_backend.cpp:88-97).  The server is the reference's
`WebSocketServer("svo", 8001, backend)` (src/app/main.cpp:208, websocketserver.cpp:35-55) on the standard library only: a
minimal RFC 6455 endpoint (text frames, ping/pong, close).  Like the reference, every call into the slam object happens on
ONE thread — the caller's: `serve_pending()` is polled between `new_image` calls (the Qt event loop of the reference does
the same interleaving, main.cpp:207-217).
"""
import base64
import hashlib
import json
import select
import socket
import struct

import numpy as np

from . import synth

_GUID = b"258EAFA5-E914-47DA-95CA-C5AB0DC85B11"


def robot_angles(pose):
    """PoseManager::get_robot_angles (pose_manager.cpp:45-59)."""
    from .cli import _rodrigues_inv
    Rx, Ry, Rz = synth._rodrigues([pose[3], 0, 0]), synth._rodrigues([0, pose[4], 0]), synth._rodrigues([0, 0, pose[5]])
    return _rodrigues_inv(Rz @ (Rx @ Ry))


def _pose_obj(pose):
    a = robot_angles(pose)
    return {"x": float(pose[0]), "y": float(pose[1]), "z": float(pose[2]), "rx": float(a[0]), "ry": float(a[1]), "rz": float(a[2])}


def keyframes_json(slam):
    """svo_slam_backend.cpp:27-66"""
    out = []
    for kf in slam.get_keyframes():
        k3, col = np.asarray(kf.kps.kps3d, np.float32), np.asarray(kf.kps.info["color"])
        out.append({"pose": _pose_obj(kf.pose),
                    "keypoints": [{"x": float(p[0]), "y": float(p[1]), "z": float(p[2])} for p in k3],
                    "colors": [{"r": int(c[0]), "g": int(c[1]), "b": int(c[2])} for c in col]})
    return json.dumps(out, separators=(",", ":"))


def pose_json(slam):
    """svo_slam_backend.cpp:68-84 (the reference dereferences an empty frame before the first image; here: zero pose)"""
    f = slam.get_frame()
    return json.dumps({"pose": _pose_obj(f.pose if f is not None else np.zeros(6))}, separators=(",", ":"))


def trajectory_json(slam):
    """svo_slam_backend.cpp:85-102"""
    t = np.asarray(slam.get_trajectory(), np.float32).reshape(-1)
    return json.dumps({"trajectory": [float(v) for v in t]}, separators=(",", ":"))


def answer(slam, resource, message):
    """text_message_received: resource name after the last '/', then the dispatch of :26-103. None = no answer."""
    url = resource.rsplit("/", 1)[-1]
    if url == "keyframes":
        return keyframes_json(slam) if message == "get" else None
    if url == "pose":
        return pose_json(slam)
    if url == "trajectory":
        return trajectory_json(slam)
    return None


# ------------------------------------------------------------------------------------------------ RFC 6455, minimal
def _frame(opcode, payload):
    n = len(payload)
    head = bytes([0x80 | opcode])
    if n < 126:
        head += bytes([n])
    elif n < 65536:
        head += bytes([126]) + struct.pack(">H", n)
    else:
        head += bytes([127]) + struct.pack(">Q", n)
    return head + payload


MAX_REQUEST_BYTES = 64 * 1024   # the viewer's requests are a few bytes; anything larger is dropped with the connection
SEND_TIMEOUT_S = 2.0            # a viewer that stops reading must not stall the tracking loop for longer


class _Client:
    def __init__(self, sock):
        self.sock, self.buf, self.resource, self.open = sock, b"", None, True

    def send(self, data):
        """sendall that never raises into the tracking loop: a viewer that went away (or stalls) just loses its connection"""
        try:
            self.sock.sendall(data)
            return True
        except OSError:      # BrokenPipeError, ConnectionResetError, socket.timeout
            self.open = False
            return False

    def handshake(self):
        if b"\r\n\r\n" not in self.buf:
            return
        head, self.buf = self.buf.split(b"\r\n\r\n", 1)
        lines = head.decode("latin1").split("\r\n")
        self.resource = lines[0].split(" ")[1] if len(lines[0].split(" ")) > 1 else "/"
        hdr = {k.strip().lower(): v.strip() for k, v in (ln.split(":", 1) for ln in lines[1:] if ":" in ln)}
        key = hdr.get("sec-websocket-key")
        if not key:
            self.send(b"HTTP/1.1 400 Bad Request\r\n\r\n")
            self.open = False
            return
        acc = base64.b64encode(hashlib.sha1(key.encode() + _GUID).digest())
        self.send(b"HTTP/1.1 101 Switching Protocols\r\nUpgrade: websocket\r\nConnection: Upgrade\r\nSec-WebSocket-Accept: " + acc + b"\r\n\r\n")

    def frames(self):
        """complete frames in the buffer -> (opcode, payload); fragmented messages are not used by the viewer"""
        while len(self.buf) >= 2:
            b0, b1 = self.buf[0], self.buf[1]
            n, off = b1 & 127, 2
            if n == 126:
                if len(self.buf) < 4:
                    return
                n, off = struct.unpack(">H", self.buf[2:4])[0], 4
            elif n == 127:
                if len(self.buf) < 10:
                    return
                n, off = struct.unpack(">Q", self.buf[2:10])[0], 10
            masked = b1 & 0x80
            if n > MAX_REQUEST_BYTES:     # never buffer what a client merely announces
                self.open = False
                self.buf = b""
                return
            if len(self.buf) < off + (4 if masked else 0) + n:
                return
            mask = self.buf[off:off + 4] if masked else None
            off += 4 if masked else 0
            payload = self.buf[off:off + n]
            if mask:
                payload = bytes(np.frombuffer(payload, np.uint8) ^ np.resize(np.frombuffer(mask, np.uint8), n)) if n else b""
            self.buf = self.buf[off + n:]
            yield b0 & 15, payload


class WebSocketServer:
    """WebSocketServer("svo", 8001, backend) of the reference (websocketserver.cpp:35-55), polled from the tracking thread."""

    def __init__(self, slam, port=8001, host="127.0.0.1"):
        """host: loopback by default — the reference's server listens on every interface (QHostAddress::Any); pass "0.0.0.0" to do
        the same (the keyframe dump is unauthenticated)"""
        self.slam = slam
        self.srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.srv.bind((host, port))
        self.srv.listen(8)
        self.srv.setblocking(False)
        self.port = self.srv.getsockname()[1]
        self.clients = []

    def serve_pending(self, timeout=0.0):
        """Accept connections and answer every message that has arrived; returns the number of answers sent."""
        sent = 0
        socks = [self.srv] + [c.sock for c in self.clients]
        ready, _, _ = select.select(socks, [], [], timeout)
        for s in ready:
            if s is self.srv:
                conn, _ = self.srv.accept()
                conn.settimeout(SEND_TIMEOUT_S)
                self.clients.append(_Client(conn))
                continue
            c = next(x for x in self.clients if x.sock is s)
            try:
                data = s.recv(65536)
            except OSError:
                data = b""
            if not data:
                c.open = False
                continue
            c.buf += data
            if len(c.buf) > 2 * MAX_REQUEST_BYTES:
                c.open = False
                continue
            if c.resource is None:
                c.handshake()
                if c.resource is None or not c.open:
                    continue
            for op, payload in c.frames():
                if op == 1:      # text
                    out = answer(self.slam, c.resource, payload.decode("utf8", "replace"))
                    if out is not None and c.send(_frame(1, out.encode("utf8"))):
                        sent += 1
                elif op == 9:    # ping
                    c.send(_frame(10, payload))
                elif op == 8:    # close
                    c.send(_frame(8, payload[:2]))
                    c.open = False
                if not c.open:
                    break
        for c in [x for x in self.clients if not x.open]:
            c.sock.close()
            self.clients.remove(c)
        return sent

    def close(self):
        for c in self.clients:
            c.sock.close()
        self.clients = []
        self.srv.close()
