"""TensorBoard scalar summaries without tensorboardX (absent from this image): a minimal event-file writer with the
`SummaryWriter.add_scalar / flush / close` surface the reference's runner uses (algo/runners/runner.py:368-423,
`runner.writers[policy_id]`).  Files are standard TFRecord-framed `Event` protos (`events.out.tfevents.*`) that TensorBoard
reads: record = uint64 length | masked crc32c(length) | payload | masked crc32c(payload); the two protobuf messages needed
(Event{wall_time, step, file_version | summary{value{tag, simple_value}}}) are encoded by hand."""
from __future__ import annotations

import os
import socket
import struct
import time

_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _len_field(field: int, payload: bytes) -> bytes:
    return bytes([(field << 3) | 2]) + _varint(len(payload)) + payload


def _event(wall_time: float, step: int = 0, file_version: str = None, tag: str = None, value: float = None) -> bytes:
    ev = b"\x09" + struct.pack("<d", wall_time) + b"\x10" + _varint(step)
    if file_version is not None:
        ev += _len_field(3, file_version.encode())
    if tag is not None:
        val = _len_field(1, tag.encode()) + b"\x15" + struct.pack("<f", float(value))
        ev += _len_field(5, _len_field(1, val))
    return ev


class SummaryWriter:
    def __init__(self, logdir: str, flush_secs: int = 30):
        os.makedirs(logdir, exist_ok=True)
        self.logdir = logdir
        name = f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}.{os.getpid()}.sfb200"
        self.path = os.path.join(logdir, name)
        self._f = open(self.path, "ab")
        self._write(_event(time.time(), 0, file_version="brain.Event:2"))
        self.flush()

    def _write(self, payload: bytes) -> None:
        header = struct.pack("<Q", len(payload))
        self._f.write(header + struct.pack("<I", _masked_crc(header)) + payload + struct.pack("<I", _masked_crc(payload)))

    def add_scalar(self, tag: str, scalar_value, global_step: int = 0, walltime: float = None) -> None:
        self._write(_event(time.time() if walltime is None else walltime, int(global_step), tag=tag, value=float(scalar_value)))

    def flush(self) -> None:
        self._f.flush()

    def close(self) -> None:
        if not self._f.closed:
            self._f.flush()
            self._f.close()


def _fields(msg: bytes):
    """[(field number, wire type, value)] of one protobuf message (varint, 64-bit, length-delimited, 32-bit)"""
    out, i = [], 0
    while i < len(msg):
        key, sh = 0, 0
        while True:
            b = msg[i]
            i += 1
            key |= (b & 0x7F) << sh
            sh += 7
            if not b & 0x80:
                break
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, sh = 0, 0
            while True:
                b = msg[i]
                i += 1
                v |= (b & 0x7F) << sh
                sh += 7
                if not b & 0x80:
                    break
        elif wt == 1:
            v = msg[i:i + 8]
            i += 8
        elif wt == 5:
            v = msg[i:i + 4]
            i += 4
        else:
            ln, sh = 0, 0
            while True:
                b = msg[i]
                i += 1
                ln |= (b & 0x7F) << sh
                sh += 7
                if not b & 0x80:
                    break
            v = msg[i:i + ln]
            i += ln
        out.append((field, wt, v))
    return out


def read_scalars(path: str):
    """[(step, tag, value)] of an event file written by SummaryWriter (verifies both checksums; used by the tests)."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        header = data[pos:pos + 8]
        (n,) = struct.unpack("<Q", header)
        assert struct.unpack("<I", data[pos + 8:pos + 12])[0] == _masked_crc(header), "length checksum"
        payload = data[pos + 12:pos + 12 + n]
        assert struct.unpack("<I", data[pos + 12 + n:pos + 16 + n])[0] == _masked_crc(payload), "payload checksum"
        pos += 16 + n
        ev = _fields(payload)
        step = next((v for f, wt, v in ev if f == 2), 0)
        for f, wt, v in ev:
            if f == 5:                                       # Event.summary
                for f2, _, value in _fields(v):              # Summary.value
                    if f2 == 1:
                        vf = {ff: vv for ff, _, vv in _fields(value)}
                        out.append((step, vf[1].decode(), struct.unpack("<f", vf[2])[0]))
    return out
