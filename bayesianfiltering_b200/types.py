"""Type aliases mirroring gaussfiltax/types.py:1-9."""
from typing import Union

import numpy as np

PRNGKey = np.ndarray
Scalar = Union[float, np.floating]
