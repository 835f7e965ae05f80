"""TREKIS-3 run on the B200 engine: `python scripts/trekis3_run.py RUN_DIR [--nmc N] [--tables-only] ...`
(under torchrun: one rank per GPU).  See trekis-3_b200/main.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trekis3_b200  # noqa: E402,F401
from trekis3_b200.main import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
