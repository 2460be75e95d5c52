#!/usr/bin/env python
"""tools/bench_monolithic.py -- the reference's monolithic bench harness
(monolithic/scripts/bench_and_plot_monolithic.sh:32-62) without its dependencies on /usr/bin/time, `column`
and matplotlib: same sweep (threads x runs of `<exe> <image> <out.png> <passes>` under OMP_NUM_THREADS), same
CSV schema (threads,avg_real_sec,std_real_sec,avg_cpu_pct,avg_mem_kb), so the reference's plots and the frontend
charts read the file unchanged.  Works for any build of the CLI: the reference's own, or the same main.c linked
against libompimg_b200 (oracle/_ref/bin/grayscale_b200), with or without OMPIMG_WORKER in the environment.

    python tools/bench_monolithic.py --exe oracle/_ref/bin/grayscale_b200 --image oracle/_ref/fixtures/test.jpg \
        --threads 1 2 4 6 --runs 10 --passes 1 --csv /tmp/monolithic_bench.csv
"""
import argparse
import math
import os
import resource
import subprocess
import tempfile
import time


def run_once(exe, image, out_png, passes, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    before = resource.getrusage(resource.RUSAGE_CHILDREN)
    t0 = time.perf_counter()
    r = subprocess.run([exe, image, out_png, str(passes)], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    real = time.perf_counter() - t0
    after = resource.getrusage(resource.RUSAGE_CHILDREN)
    if r.returncode != 0:
        raise SystemExit(f"{exe} failed ({r.returncode}): {r.stderr.strip()}")
    cpu = (after.ru_utime - before.ru_utime) + (after.ru_stime - before.ru_stime)
    return real, 100.0 * cpu / real if real > 0 else 0.0, after.ru_maxrss          # ru_maxrss of children: high-water mark, KiB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exe", required=True)
    ap.add_argument("--image", required=True)
    ap.add_argument("--threads", type=int, nargs="+", default=[1, 2, 3, 4, os.cpu_count() or 1])
    ap.add_argument("--runs", type=int, default=10)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--csv", default="monolithic_bench.csv")
    a = ap.parse_args()
    rows = ["threads,avg_real_sec,std_real_sec,avg_cpu_pct,avg_mem_kb"]
    with tempfile.TemporaryDirectory() as tmp:
        for t in a.threads:
            reals, cpus, mems = [], [], []
            for run in range(a.runs):
                real, cpu, mem = run_once(a.exe, a.image, os.path.join(tmp, f"output_{t}_{run}.png"), a.passes, t)
                reals.append(real); cpus.append(cpu); mems.append(mem)
            avg = sum(reals) / len(reals)
            std = math.sqrt(max(0.0, sum(x * x for x in reals) / len(reals) - avg * avg))
            rows.append("%d,%.5f,%.5f,%.1f,%.0f" % (t, avg, std, sum(cpus) / len(cpus), sum(mems) / len(mems)))
            print(f">> OMP_NUM_THREADS={t}  (x{a.runs} run, kernel x{a.passes}): {avg:.4f} s +- {std:.4f}")
    with open(a.csv, "w") as f:
        f.write("\n".join(rows) + "\n")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
