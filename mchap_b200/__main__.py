"""``python -m mchap_b200 assemble|call|call-exact ...`` (see mchap_b200/application/cli.py)."""
import sys

from .application.cli import main

if __name__ == "__main__":
    sys.exit(main(["mchap"] + sys.argv[1:]))
