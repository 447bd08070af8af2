import sys

from . import FLAVOURS, path

if len(sys.argv) != 2 or sys.argv[1] not in FLAVOURS:
    sys.exit("usage: python -m sparsefactorization_b200.compat {%s}" % "|".join(FLAVOURS))
print(path(sys.argv[1]))
