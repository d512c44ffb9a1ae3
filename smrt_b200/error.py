"""Error and warning types, mirroring reference ``smrt/core/error.py:6-29``."""

import warnings


class SMRTError(Exception):
    """Error raised by the model (same name and meaning as the reference's)."""


class SMRTWarning(Warning):
    """Warning category for numerical or physical concerns that do not stop the computation."""


def smrt_warn(message, category=SMRTWarning, stacklevel=2):
    warnings.warn(message, category, stacklevel=stacklevel + 1)
