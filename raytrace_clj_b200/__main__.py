import sys

from .core import main

sys.exit(main())
