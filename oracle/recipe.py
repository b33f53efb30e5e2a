"""ORACLE — test infrastructure.  The synthetic weight / mask / frame recipe lives with the product
(drnb200.synthetic) because bench.py needs it without touching oracle/; re-exported here for the tests."""
from drnb200.synthetic import *  # noqa: F401,F403
from drnb200.synthetic import _gen  # noqa: F401
