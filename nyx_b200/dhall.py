"""Reader for the DATA subset of Dhall the reference writes and reads its configuration in.

The reference (de)serialises `PropagatorConfig`, the sequence's `propagators` map, the almanac file list … through `serde_dhall`
(dynamics/sequence/config.rs:96-169; data/02_config/*.dhall).  What `serde_dhall` emits is Dhall in normal form: records, lists,
`Some x` / `None T`, union literals `< A | B : T >.B payload`, numbers, booleans and text — no imports, functions or `let`.  That
subset is all this reader accepts; anything else raises `DhallError` (resolve it with the `dhall` tool first).

Mapping to Python: record -> dict, list -> list, `Some x` -> x, `None T` -> None, union alternative without payload -> its name
(str), with payload -> {name: payload}, Natural / Integer -> int, Double -> float, Bool -> bool, Text -> str.
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Any, List, Tuple, Union


class DhallError(ValueError):
    pass


_TOKEN = re.compile(r"""
    (?P<ws>\s+|--[^\n]*|\{-.*?-\})
  | (?P<text>"(?:[^"\\]|\\.)*")
  | (?P<num>[+-]?(?:\d+\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+|\d+|Infinity)|NaN)
  | (?P<ident>`[^`]+`|[A-Za-z_][A-Za-z0-9_/\-]*)
  | (?P<punct>[{}\[\]<>(),=:|.])
""", re.X | re.S)

_ESCAPES = {'"': '"', "\\": "\\", "/": "/", "b": "\b", "f": "\f", "n": "\n", "r": "\r", "t": "\t", "$": "$"}


def _tokenize(src: str) -> List[Tuple[str, str]]:
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            line = src.count("\n", 0, pos) + 1
            raise DhallError(f"line {line}: unexpected character {src[pos]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group(kind)))
    out.append(("eof", ""))
    return out


def _unescape(lit: str) -> str:
    body, out, i = lit[1:-1], [], 0
    while i < len(body):
        c = body[i]
        if c == "\\":
            nxt = body[i + 1]
            if nxt == "u":
                out.append(chr(int(body[i + 2:i + 6], 16)))
                i += 6
                continue
            if nxt not in _ESCAPES:
                raise DhallError(f"unknown escape \\{nxt}")
            out.append(_ESCAPES[nxt])
            i += 2
        else:
            if c == "$" and body[i + 1:i + 2] == "{":
                raise DhallError("text interpolation is outside the data subset")
            out.append(c)
            i += 1
    return "".join(out)


class _Parser:
    def __init__(self, src: str):
        self.toks = _tokenize(src)
        self.i = 0

    def peek(self) -> Tuple[str, str]:
        return self.toks[self.i]

    def next(self) -> Tuple[str, str]:
        t = self.toks[self.i]
        self.i += 1
        return t

    def expect(self, value: str) -> None:
        kind, v = self.next()
        if v != value:
            raise DhallError(f"expected {value!r}, found {v or kind!r}")

    def label(self) -> str:
        kind, v = self.next()
        if kind != "ident":
            raise DhallError(f"expected a label, found {v or kind!r}")
        return v.strip("`")

    # ---- types (only ever skipped: after `None`, inside union declarations, after `[] :`)
    def _type_atom_ahead(self) -> bool:
        kind, v = self.peek()
        return kind == "ident" or v in ("{", "<", "(")

    def skip_type(self) -> None:
        if not self._type_atom_ahead():
            raise DhallError(f"expected a type, found {self.peek()[1]!r}")
        while self._type_atom_ahead():   # application by juxtaposition: `List { .. }`, `Optional Double`
            kind, v = self.next()
            if v == "{":
                if self.peek()[1] == "}":
                    self.next()
                    continue
                while True:
                    self.label()
                    self.expect(":")
                    self.skip_type()
                    if self.peek()[1] == ",":
                        self.next()
                        continue
                    self.expect("}")
                    break
            elif v == "<":
                self.union_alternatives()
            elif v == "(":
                self.skip_type()
                self.expect(")")

    def union_alternatives(self) -> dict:
        """after '<': {alternative: has_payload}, consumes the closing '>'"""
        alts = {}
        if self.peek()[1] == ">":
            self.next()
            return alts
        while True:
            name = self.label()
            has_type = self.peek()[1] == ":"
            if has_type:
                self.next()
                self.skip_type()
            alts[name] = has_type
            kind, v = self.next()
            if v == ">":
                return alts
            if v != "|":
                raise DhallError(f"expected '|' or '>' in a union type, found {v or kind!r}")

    # ---- values
    def value(self) -> Any:
        kind, v = self.next()
        if kind == "text":
            return _unescape(v)
        if kind == "num":
            if v in ("NaN", "Infinity", "+Infinity", "-Infinity"):
                return float(v.replace("Infinity", "inf").replace("NaN", "nan"))
            return float(v) if ("." in v or "e" in v or "E" in v) else int(v)
        if kind == "ident":
            if v == "True":
                return True
            if v == "False":
                return False
            if v == "Some":
                return self.value()
            if v == "None":
                self.skip_type()
                return None
            raise DhallError(f"{v!r}: only records, lists, Some/None, unions and literals belong to the data subset")
        if v == "{":
            return self.record()
        if v == "[":
            return self.list()
        if v == "<":
            alts = self.union_alternatives()
            self.expect(".")
            name = self.label()
            if name not in alts:
                raise DhallError(f"{name!r} is not an alternative of the union")
            return {name: self.value()} if alts[name] else name
        if v == "(":
            inner = self.value()
            self.expect(")")
            return inner
        raise DhallError(f"unexpected {v or kind!r}")

    def record(self) -> dict:
        out: dict = {}
        if self.peek()[1] == "=":      # `{=}`: the empty record
            self.next()
            self.expect("}")
            return out
        if self.peek()[1] == "}":
            self.next()
            return out
        while True:
            key = self.label()
            self.expect("=")
            out[key] = self.value()
            kind, v = self.next()
            if v == "}":
                return out
            if v != ",":
                raise DhallError(f"expected ',' or '}}' after field {key!r}, found {v or kind!r}")

    def list(self) -> list:
        out: list = []
        if self.peek()[1] == "]":
            self.next()
        else:
            while True:
                out.append(self.value())
                kind, v = self.next()
                if v == "]":
                    break
                if v != ",":
                    raise DhallError(f"expected ',' or ']' in a list, found {v or kind!r}")
        if self.peek()[1] == ":":      # `[] : List T`
            self.next()
            self.skip_type()
        return out


def loads(src: str) -> Any:
    p = _Parser(src)
    v = p.value()
    if p.peek()[0] != "eof":
        raise DhallError(f"trailing input at {p.peek()[1]!r}")
    return v


def load(path: Union[str, Path]) -> Any:
    return loads(Path(path).read_text())


def pairs_to_dict(v: Any) -> Any:
    """serde_dhall writes maps as lists of `{ _1 = key, _2 = value }` (or `{ mapKey, mapValue }`): turn such a list into a dict."""
    if isinstance(v, list) and v and all(isinstance(e, dict) and (set(e) == {"_1", "_2"} or set(e) == {"mapKey", "mapValue"}) for e in v):
        return {(e["_1"] if "_1" in e else e["mapKey"]): (e["_2"] if "_2" in e else e["mapValue"]) for e in v}
    return v
