"""Legacy Gaussian-sum filter with the call surface of gaussfiltax/gausssumfilt.py:11-78, running on libbfilt.so.

`GaussSumFilt(ssm, M).run(ys, m0, P0)` -> (means [T+1,dx,M], covs [T+1,dx,dx,M], weights [T+1,M], point_est [T,dx]).
Reference quirks kept (they are the class's behaviour): the prediction adds the covariance itself instead of Q
(`cov + F cov F^T`, gausssumfilt.py:59), the components start at m0 + N(0, I) (:47), the initial condition sits at
index T, and the point estimate sums over every axis (:76).
"""
from . import _lib
from .gaussfilt import EKF, GaussSumFilt as _BankOfEKFs


class GaussSumFilt(_BankOfEKFs):
    _flags = _lib.GSF_PREDICT_FIRST | _lib.GSF_NO_JITTER | _lib.GSF_COV_PLUS_P

    def __init__(self, ssm, M):
        _BankOfEKFs.__init__(self, EKF(ssm, order=1), M)
        self.f, self.g, self.Q, self.R = ssm.f, ssm.g, ssm.Q, ssm.R

    def __str__(self):
        return 'GSF'
