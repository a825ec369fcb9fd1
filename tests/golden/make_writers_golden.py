"""Regenerates tests/golden/writers.json: digests of the files the REFERENCE's writers produce for the call sequences of
tests/cpp/track_files_test.cpp (src/gpx.cpp, src/ptu.cpp, SD/io/{csv,gpx,kml}.c compiled unmodified into
oracle/_ref/libwriters_ref.so by `make -C oracle writers`).  Needs /root/reference at build time; the committed JSON lets
tests/test_track_files.py pin host/track_files.hpp where the compiled reference is not available.

    python tests/golden/make_writers_golden.py
"""
import hashlib
import json
import os
import subprocess
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
LIB = os.path.join(ROOT, "oracle", "_ref", "libwriters_ref.so")
EXE = os.path.join(ROOT, "build", "track_files_test")
ROUNDS = (1, 2, 4, 7, 12)                     # a run of k rounds leaves the files of round k-1 (seeds 1000/2000/3000 + k-1)
FILES = ("w.gpx", "ptu.csv", "c.csv", "c.gpx", "c.kml", "l.kml-live.kml")      # l.kml names the live file by path: left out


def digests(prefix, lib):
    out = {}
    for k in ROUNDS:
        with tempfile.TemporaryDirectory() as d:
            subprocess.run([EXE, lib, d, str(k)], check=True, capture_output=True)
            for f in FILES:
                b = open(os.path.join(d, prefix + f), "rb").read()
                out[f"{k}/{f}"] = {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}
    return out


if __name__ == "__main__":
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "cli", "writers"], check=True)
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "cpp", "track_files_test.cpp"),
                    "-o", EXE, "-ldl"], check=True)
    ref = digests("ref_", LIB)
    with open(os.path.join(ROOT, "tests", "golden", "writers.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_writers_golden.py", "source": "reference writers (oracle/_ref/libwriters_ref.so)",
                   "files": ref}, f, indent=1, sort_keys=True)
    print(f"{len(ref)} digests written")
