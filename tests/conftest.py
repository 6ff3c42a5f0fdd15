import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box)')


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def load_golden(path):
    with np.load(path, allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    for k in ('kind', 'layout'):
        d[k] = str(d[k])
    d['layout'] = json.loads(d['layout'])
    for k in ('N', 'ndec', 'ncons'):
        d[k] = int(d[k])
    for k in ('dt', 'obj_factor', 'f'):
        d[k] = float(d[k])
    d['dims'] = tuple(int(v) for v in d['dims'])
    d['name'] = os.path.basename(path)[:-4]
    return d


@pytest.fixture(params=golden_files(),
                ids=lambda p: os.path.basename(p)[:-4])
def golden(request):
    return load_golden(request.param)
