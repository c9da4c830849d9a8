"""Import shim, TEST INFRASTRUCTURE ONLY (never on the product path).

The reference's scheduler imports `diffusers` for config plumbing only
(/root/reference/src/models/lcm_scheduler.py:23-24,34,53); `diffusers` is not
installed in the build container.  This stub provides just enough
(`SchedulerMixin`, `ConfigMixin`, `register_to_config`) for the unmodified
reference to import, so that `tests/golden/make_golden.py` can run it and pin
the oracle.  No arithmetic lives here.
"""
from .configuration_utils import ConfigMixin, register_to_config  # noqa: F401


class SchedulerMixin:  # marker base class only
    pass
