"""Row filters for `.where(...)`: a small SQL-predicate evaluator over the table's non-vector columns.

The reference hands the filter string to DataFusion (python/python/lancedb/query.py:1864-1892 `where`,
rust/lancedb/src/query.rs:489-507 prefilter / postfilter).  Filter evaluation is not part of the GPU hot
path; what the hot path consumes is its *result*: the allow-list of row ids, as a bitmap
(`lgpu_search_filtered`, include/lancedb_b200.h).  This module produces that mask on the host with
pyarrow.compute for the predicate forms the reference's own tests use (`id = 2`, `id >= 1`, `b < 10`,
AND / OR / NOT, IN, BETWEEN, IS [NOT] NULL, LIKE); anything else raises ValueError, as an unparsable
filter does in the reference.  SQL three-valued logic: a NULL predicate excludes the row.
"""
from __future__ import annotations

import re
from typing import List, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

_TOKEN = re.compile(r"""
    \s*(?:
      (?P<num>\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)
    | (?P<str>'(?:[^']|'')*')
    | (?P<qid>`[^`]+`|"[^"]+")
    | (?P<id>[A-Za-z_][A-Za-z_0-9.]*)
    | (?P<op><=|>=|<>|!=|==|=|<|>|\(|\)|,|-|\+)
    )""", re.X)

_KEYWORDS = {"AND", "OR", "NOT", "IN", "IS", "NULL", "BETWEEN", "LIKE", "TRUE", "FALSE"}


def _tokenize(text: str) -> List[Tuple[str, object]]:
    out, pos = [], 0
    text = text.strip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise ValueError(f"cannot parse filter near {text[pos:pos + 20]!r}")
        pos = m.end()
        if m.group("num") is not None:
            t = m.group("num")
            out.append(("lit", float(t) if any(c in t for c in ".eE") else int(t)))
        elif m.group("str") is not None:
            out.append(("lit", m.group("str")[1:-1].replace("''", "'")))
        elif m.group("qid") is not None:
            out.append(("col", m.group("qid")[1:-1]))
        elif m.group("id") is not None:
            w = m.group("id")
            if w.upper() in _KEYWORDS:
                out.append(("kw", w.upper()))
            else:
                out.append(("col", w))
        else:
            out.append(("op", m.group("op")))
    return out


class _Parser:
    def __init__(self, tokens, table: pa.Table):
        self.t, self.i, self.table = tokens, 0, table

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, kind=None, val=None):
        k, v = self.peek()
        if k is None or (kind and k != kind) or (val is not None and v != val):
            raise ValueError(f"filter syntax error at token {self.i}: expected {val or kind}, got {v!r}")
        self.i += 1
        return v

    def accept(self, kind, val):
        if self.peek() == (kind, val):
            self.i += 1
            return True
        return False

    # ---- boolean structure ----
    def parse(self):
        e = self.or_()
        if self.i != len(self.t):
            raise ValueError(f"filter syntax error: unexpected {self.peek()[1]!r}")
        return e

    def or_(self):
        e = self.and_()
        while self.accept("kw", "OR"):
            e = pc.or_kleene(e, self.and_())
        return e

    def and_(self):
        e = self.not_()
        while self.accept("kw", "AND"):
            e = pc.and_kleene(e, self.not_())
        return e

    def not_(self):
        if self.accept("kw", "NOT"):
            return pc.invert(self.not_())
        return self.predicate()

    # ---- predicates ----
    def operand(self):
        k, v = self.peek()
        if k == "op" and v in "-+":
            self.i += 1
            x = self.operand()
            return pc.negate(x) if v == "-" else x
        if k == "lit":
            self.i += 1
            return pa.scalar(v)
        if k == "kw" and v in ("TRUE", "FALSE"):
            self.i += 1
            return pa.scalar(v == "TRUE")
        if k == "col":
            self.i += 1
            if v not in self.table.column_names:
                raise ValueError(f"filter refers to unknown column {v!r}")
            return self.table.column(v)
        raise ValueError(f"filter syntax error at token {self.i}: unexpected {v!r}")

    def predicate(self):
        if self.peek() == ("op", "("):
            # parenthesised boolean expression (operands never start with "(" in this grammar)
            self.i += 1
            e = self.or_()
            self.take("op", ")")
            return e
        lhs = self.operand()
        k, v = self.peek()
        if k == "op" and v in ("=", "==", "!=", "<>", "<", "<=", ">", ">="):
            self.i += 1
            rhs = self.operand()
            fn = {"=": pc.equal, "==": pc.equal, "!=": pc.not_equal, "<>": pc.not_equal, "<": pc.less,
                  "<=": pc.less_equal, ">": pc.greater, ">=": pc.greater_equal}[v]
            return fn(lhs, rhs)
        if k == "kw" and v == "IS":
            self.i += 1
            neg = self.accept("kw", "NOT")
            self.take("kw", "NULL")
            e = pc.is_null(lhs)
            return pc.invert(e) if neg else e
        neg = False
        if k == "kw" and v == "NOT":
            self.i += 1
            neg = True
            k, v = self.peek()
        if k == "kw" and v == "IN":
            self.i += 1
            self.take("op", "(")
            vals = [self.take("lit")]
            while self.accept("op", ","):
                vals.append(self.take("lit"))
            self.take("op", ")")
            e = pc.is_in(lhs, value_set=pa.array(vals))
            e = pc.if_else(pc.is_null(lhs), pa.scalar(None, pa.bool_()), e)
        elif k == "kw" and v == "BETWEEN":
            self.i += 1
            lo = self.operand()
            self.take("kw", "AND")
            hi = self.operand()
            e = pc.and_kleene(pc.greater_equal(lhs, lo), pc.less_equal(lhs, hi))
        elif k == "kw" and v == "LIKE":
            self.i += 1
            e = pc.match_like(lhs, self.take("lit"))
        else:
            if neg:
                raise ValueError("filter syntax error after NOT")
            # a bare boolean column / literal
            return lhs
        return pc.invert(e) if neg else e


def evaluate(table: pa.Table, where: str) -> np.ndarray:
    """Boolean mask over the table's rows (row id = row position) for the SQL predicate `where`."""
    if not isinstance(where, str) or not where.strip():
        raise ValueError("filter must be a non-empty SQL predicate string")
    res = _Parser(_tokenize(where), table).parse()
    if isinstance(res, pa.Scalar):
        return np.full(table.num_rows, bool(res.as_py()) if res.is_valid else False)
    if isinstance(res, pa.ChunkedArray):
        res = res.combine_chunks() if res.num_chunks != 1 else res.chunk(0)
    if not pa.types.is_boolean(res.type):
        raise ValueError("filter does not evaluate to a boolean")
    return np.asarray(pc.fill_null(res, False).to_numpy(zero_copy_only=False), bool)


def combine(existing, new: str) -> str:
    """`where` called twice ANDs the filters (python/python/lancedb/query.py:125-143)."""
    if existing is None:
        return new
    return f"({existing}) AND ({new})"
