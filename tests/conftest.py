import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "tests", "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# fixtures of the reference's own test-suite (test/runtests.jl:8-22)
S1 = "(D:18.03,(C:12.06,(B:7.06,A:7.06):4.99):5.97);"
DF1 = (["A", "B", "C", "D"], np.array([[2, 2, 3, 4], [3, 5, 0, 1]], dtype=np.int64))
SHOULDBE1 = [-np.inf, -13.0321, -10.2906, -8.96844, -8.41311, -8.38041, -8.78481, -9.5921, -10.8016, -12.4482, -14.6268, -17.607]
SHOULDBE2 = [-np.inf, -12.6372, -10.112, -8.97727, -8.59388, -8.72674, -9.29184, -10.2568, -11.6216, -13.4219, -15.753, -18.8843]
SHOULDBE3 = [-1296.1763022581217, -1408.4459849625102, -1172.356609929616, -1197.7833487345943, -3849.887791929266,
             -1348.7468163817248, -1486.4818702521393, -2083.8730465634676, -2351.3561862674096, -1364.8837334043392]
LAM3 = [0.56, 0.97, 0.12, 0.21, 4.23, 0.12, 0.77, 2.01, 2.99, 0.44]
MU3 = [0.64, 0.36, 0.25, 0.57, 0.42, 0.49, 0.61, 0.69, 3.92, 0.53]
ETA3 = [0.55, 0.56, 0.93, 0.57, 0.59, 0.06, 0.15, 0.41, 0.96, 0.12]
GRAD_GOLDEN = [0.0, 0.21654, -4.29457, 0.28282, -0.07077, -1.77586, -2.57598, 6.25627, -2.82879, 1.81408, 3.84539,
               19.34362, -0.47348, -1.39255, -2.83181, 27.41048, -1.44903, -2.66647, 3.29678, 0.0, -0.08539, 2.6401,
               -0.37875, 0.12269, -1.05525, -1.48215, -1.38917, 7.49706, 0.25451, -2.1514, -10.16407, 0.06406, 1.94643,
               11.01334, -1.90087, 1.26579, 22.01815, -3.40453, 49.2567]


def read_nw(name):
    with open(os.path.join(DATA, name)) as fh:
        return fh.readline().strip()


def read_csv(name):
    path = os.path.join(DATA, name)
    with open(path) as fh:
        names = fh.readline().strip().split(",")
    counts = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.int64, ndmin=2)
    return names, counts


@pytest.fixture(scope="session")
def plants1():
    return read_nw("plants1.nw"), read_csv("plants1-100.csv")


@pytest.fixture(scope="session")
def plants1c():
    return read_nw("plants1c.nw"), read_csv("plants1-100.csv")


@pytest.fixture(scope="session")
def dicots():
    return read_nw("9dicots.nw"), read_csv("9dicots-f01-25.csv")


@pytest.fixture(scope="session")
def monocots():
    return read_nw("12monocots.nw"), read_csv("12monocots-f01-250.csv")
