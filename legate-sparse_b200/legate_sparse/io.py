"""mmread — MatrixMarket coordinate reader → csr_array
(reference legate_sparse/io.py:26-55 + src/sparse/io/mtx_to_coo.cc:31-143).

Same subset as the reference: ``matrix coordinate`` with field real / pattern / integer and
symmetry general / symmetric; values are float64; symmetric off-diagonal entries are
mirrored right after the entry they come from; the COO triplets then go through the COO
constructor (stable sort by row, duplicates kept).  Host-only, like upstream (single task).
"""
import numpy

from .csr import csr_array


def _read_mtx_to_coo(source):
    with open(source, "r") as f:
        header = f.readline().split()
        if len(header) < 5 or header[0] != "%%MatrixMarket":
            raise ValueError("Unknown header of MatrixMarket")
        _, mtype, fmt, field, symmetry = header[:5]
        if mtype != "matrix":
            raise ValueError("must have type matrix")
        if fmt != "coordinate":
            raise ValueError("must be coordinate")
        field = field.lower()
        if field not in ("real", "pattern", "integer"):
            raise ValueError("unknown field")
        symmetry = symmetry.lower()
        if symmetry not in ("symmetric", "general"):
            raise ValueError("unknown symmetry")
        symmetric = symmetry == "symmetric"
        line = f.readline()
        while line and line.lstrip().startswith("%"):
            line = f.readline()
        dims = line.split()
        m, n, lines = int(dims[0]), int(dims[1]), int(dims[2])
        cap = lines * 2 if symmetric else lines
        rows = numpy.empty(cap, dtype=numpy.int64)
        cols = numpy.empty(cap, dtype=numpy.int64)
        vals = numpy.empty(cap, dtype=numpy.float64)
        idx = 0
        for line in f:
            tok = line.split()
            if not tok:
                continue
            cx, cy = int(tok[0]), int(tok[1])
            if field == "pattern":
                v = 1.0
            elif field == "integer":
                v = float(int(tok[2]))
            else:
                v = float(tok[2])
            rows[idx], cols[idx], vals[idx] = cx - 1, cy - 1, v
            idx += 1
            if symmetric and cx != cy:
                rows[idx], cols[idx], vals[idx] = cy - 1, cx - 1, v
                idx += 1
    return m, n, rows[:idx], cols[:idx], vals[:idx]


def mmread(source):
    m, n, rows, cols, vals = _read_mtx_to_coo(source)
    return csr_array((vals, (rows, cols)), shape=(m, n))
