#!/usr/bin/env python3
"""Drop-in command: same flags and output as the reference script of this name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from well_duplicates_b200.prepare_cli import main  # noqa: E402

main()
