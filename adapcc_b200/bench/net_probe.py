"""Host-network bandwidth / latency time series — the motivation data the reference collected with
iperf and ping on AWS / Huawei clouds (/root/reference/cloud/band_profile.py, latency_profile.py,
cloud/trace/*). No iperf dependency: a TCP stream for `duration` seconds (bandwidth, Gb/s) and
64-byte echo round trips (latency, ms), one sample per interval, appended to a trace file.

    python -m adapcc_b200.bench.net_probe --serve                 # on the target host
    python -m adapcc_b200.bench.net_probe --host 10.0.0.2 --samples 20 --out trace.txt
"""
import argparse
import socket
import threading
import time


def serve(port: int):
    srv = socket.socket()
    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    srv.bind(("0.0.0.0", port))
    srv.listen(8)
    while True:
        c, _ = srv.accept()
        threading.Thread(target=_handle, args=(c,), daemon=True).start()


def _handle(c: socket.socket):
    with c:
        mode = c.recv(1)
        if mode == b"L":
            while True:
                d = c.recv(64)
                if not d:
                    return
                c.sendall(d)
        else:
            while c.recv(1 << 20):
                pass


def bandwidth(host, port, duration=1.0) -> float:
    buf = b"\0" * (1 << 20)
    with socket.create_connection((host, port)) as s:
        s.sendall(b"B")
        t0, sent = time.time(), 0
        while time.time() - t0 < duration:
            s.sendall(buf)
            sent += len(buf)
        return sent * 8 / (time.time() - t0) / 1e9


def latency(host, port, n=50) -> float:
    with socket.create_connection((host, port)) as s:
        s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        s.sendall(b"L")
        msg = b"x" * 64
        t0 = time.time()
        for _ in range(n):
            s.sendall(msg)
            got = 0
            while got < 64:
                got += len(s.recv(64 - got))
        return (time.time() - t0) / n * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--serve", action="store_true")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=5201)
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--interval", type=float, default=1.0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    if a.serve:
        return serve(a.port)
    rows = []
    for _ in range(a.samples):
        rows.append((time.time(), bandwidth(a.host, a.port, a.interval), latency(a.host, a.port)))
        print("bandwidth %.2f Gb/s  latency %.3f ms" % rows[-1][1:], flush=True)
    if a.out:
        with open(a.out, "a") as f:
            f.writelines("%.3f %.4f %.4f\n" % r for r in rows)


if __name__ == "__main__":
    main()
