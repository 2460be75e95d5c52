#!/usr/bin/env python
"""tools/bench_service.py -- the reference's service benchmarks without Flask / RabbitMQ / MinIO (none is in this image):

  * microservices/grayscale/scripts/bench_grayscale_service.sh:19,37 -- CSV `threads,avg_request_sec,std_request_sec,
    avg_service_sec,std_service_sec`: per thread count, RUNS requests; "service" is what app.py:32-34 times (the CLI
    subprocess) and reports in X-Elapsed, "request" is what test_client.py:37-39 times around the POST.  Here a request
    is everything app.py does between receiving the upload and sending the PNG (app.py:18-42: save the upload into a
    temporary directory, exec `bin/grayscale in out.png [passes]` with OMP_NUM_THREADS, read the PNG back); only the
    HTTP hop itself is missing.
  * event-driven/grayscale_service/app.py:54-84 -- the `times` dict of the AMQP reply, {"<threads>": mean seconds over
    `repeat` runs of the subprocess}: printed as JSON with --times-json (the frontend charts read exactly this shape,
    frontend/app.py:239-260).

Works with any build of the CLI: the reference's own (oracle/_ref/bin/grayscale_strict) or the same main.c linked against
libompimg_b200 (oracle/_ref/bin/grayscale_b200), with or without OMPIMG_WORKER=<socket> in the environment.

    python tools/bench_service.py --exe oracle/_ref/bin/grayscale_b200 --image oracle/_ref/fixtures/test.jpg \
        --threads 1 2 4 6 --runs 5 --passes 1000 --csv /tmp/service_bench.csv
"""
import argparse
import json
import math
import os
import shutil
import subprocess
import tempfile
import time


def one_request(exe, image, passes, threads):
    """app.py:18-42 without the socket: returns (request seconds, service seconds, png bytes)"""
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmpdir:
        in_path = os.path.join(tmpdir, os.path.basename(image))
        out_path = os.path.join(tmpdir, "out.png")
        shutil.copyfile(image, in_path)                           # img_file.save(in_path)
        cmd = [exe, in_path, out_path] + ([str(passes)] if passes else [])
        env = os.environ.copy()
        if threads:
            env["OMP_NUM_THREADS"] = str(threads)
        start = time.time()
        r = subprocess.run(cmd, capture_output=True, env=env)
        service = time.time() - start
        if r.returncode != 0:
            raise SystemExit(f"processing failed (HTTP 500 in the service): {r.stderr.decode()[-500:]}")
        with open(out_path, "rb") as f:
            data = f.read()                                       # send_file(out_path)
    return time.time() - t0, service, data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exe", required=True)
    ap.add_argument("--image", required=True)
    ap.add_argument("--threads", type=int, nargs="+", default=[1])
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--csv", default="service_bench.csv")
    ap.add_argument("--times-json", action="store_true", help="also print the event-driven worker's `times` dict")
    a = ap.parse_args()
    rows = ["threads,avg_request_sec,std_request_sec,avg_service_sec,std_service_sec"]
    times = {}
    for t in a.threads:
        print(f">> threads={t}  (x{a.runs} runs)")
        req, svc = [], []
        for _ in range(a.runs):
            r, s, _ = one_request(a.exe, a.image, a.passes, t)
            req.append(r); svc.append(s)
        avg_r, avg_s = sum(req) / len(req), sum(svc) / len(svc)
        std = lambda xs, m: math.sqrt(max(0.0, sum(x * x for x in xs) / len(xs) - m * m))
        rows.append("%d,%.5f,%.5f,%.5f,%.5f" % (t, avg_r, std(req, avg_r), avg_s, std(svc, avg_s)))
        times[str(t)] = avg_s
    with open(a.csv, "w") as f:
        f.write("\n".join(rows) + "\n")
    print("\n== Results ==")
    print("\n".join(rows))
    if a.times_json:
        print(json.dumps({"times": times, "passes": a.passes}))


if __name__ == "__main__":
    main()
