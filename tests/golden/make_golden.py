#!/usr/bin/env python3
"""Collect the reference's own test fixtures for the hot path into tests/golden/ref/.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py

Only DATA files are copied (FASTA/FASTQ inputs, prebuilt .skf files, golden stdout) --
never reference source code.  They are the known-answer vectors that pin the oracle
(SURVEY.md section 8c): tests/test_files_in/* and tests/test_results_correct/* of
bacpop/ska.rust v0.5.2.  A manifest with sha256 of every file is written beside them so
that the GPU box (which has no /root/reference) can verify it runs the same vectors.
"""
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("SKA_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref")

# skalo goldens are out of scope; map / weed / merge / delete fixtures pin the section 8(f) rows
SKIP_PREFIX = ("test_skalo",)


def main():
    src_in = os.path.join(REF, "tests", "test_files_in")
    src_ok = os.path.join(REF, "tests", "test_results_correct")
    if not os.path.isdir(src_in):
        sys.exit(f"reference not found at {REF}")
    manifest = {}
    for sub, src in (("in", src_in), ("correct", src_ok)):
        dst = os.path.join(OUT, sub)
        os.makedirs(dst, exist_ok=True)
        for root, _dirs, files in os.walk(src):
            for f in sorted(files):
                if f.startswith(SKIP_PREFIX):
                    continue
                rel = os.path.relpath(os.path.join(root, f), src)
                d = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), d)
                os.chmod(d, 0o644)
                with open(d, "rb") as fh:
                    manifest[f"{sub}/{rel}"] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "bacpop/ska.rust v0.5.2 tests/", "files": manifest}, fh, indent=1, sort_keys=True)
    print(f"copied {len(manifest)} fixture files into {OUT}")


if __name__ == "__main__":
    main()
