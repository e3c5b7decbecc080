"""
``import diffxpy_b200.api as de`` -- the namespace facade of /root/reference/diffxpy/api/__init__.py:1-9,
restricted to the accelerated hot path: ``de.test``, ``de.stats``, ``de.utils`` (``de.fit`` / ``de.enrich`` of
the reference are outside the scope of this build).
"""
from . import test
from . import stats
from . import utils
