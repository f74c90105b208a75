import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import warnings
    warnings.filterwarnings('ignore', category=SyntaxWarning)      # the reference's own docstrings
