"""Development only: run bench.py (or pytest) against a variant build of the library.
    python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_X.so [bench.py arguments]
    python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_X.so pytest [pytest arguments]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import icar_b200

icar_b200.LIB_PATH = os.path.abspath(sys.argv[1])
if len(sys.argv) > 2 and sys.argv[2] == "pytest":
    import pytest
    sys.exit(pytest.main(sys.argv[3:]))
sys.argv = ["bench.py"] + sys.argv[2:]
import bench

bench.main()
