"""Lenient XML reader/writer for the three AdapCC schemas (strategy ``<trees>``, logical
graph ``<graph>``, detect ``<cpu>``).

The reference parses these with tinyxml2 in C++ and xmltodict in Python
(/root/reference/csrc/allreduce.cu:86-104, /root/reference/commu.py:207-244). Some shipped
strategy files are not well-formed XML (``id='1'ip='10.28.1.30'`` with no separating space,
/root/reference/strategy/4.xml:3) and only tinyxml2 accepts them; ``xml.etree`` does not. This
reader accepts that dialect. The native runtime has an equivalent C++ reader
(csrc/schedule.cpp) and tests cross-check the two.
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional


class XmlError(ValueError):
    pass


@dataclass
class Node:
    tag: str
    attrs: Dict[str, str] = field(default_factory=dict)
    children: List["Node"] = field(default_factory=list)

    def find_all(self, tag: str) -> List["Node"]:
        return [c for c in self.children if c.tag == tag]

    def find(self, tag: str) -> Optional["Node"]:
        for c in self.children:
            if c.tag == tag:
                return c
        return None

    def iter(self) -> Iterator["Node"]:
        yield self
        for c in self.children:
            yield from c.iter()

    def get(self, key: str, default: Optional[str] = None) -> Optional[str]:
        return self.attrs.get(key, default)


_NAME_EXTRA = set("_-:.")


def _is_name(ch: str) -> bool:
    return ch.isalnum() or ch in _NAME_EXTRA


class _Cursor:
    def __init__(self, text: str):
        self.s = text
        self.i = 0

    def eof(self) -> bool:
        return self.i >= len(self.s)

    def starts(self, lit: str) -> bool:
        return self.s.startswith(lit, self.i)

    def skip_ws(self) -> None:
        while not self.eof() and self.s[self.i].isspace():
            self.i += 1

    def skip_misc(self) -> None:
        """Skip text, comments, <?...?> and <!...> up to the next element tag."""
        while not self.eof():
            if self.s[self.i] != "<":
                self.i += 1
            elif self.starts("<!--"):
                e = self.s.find("-->", self.i + 4)
                if e < 0:
                    raise XmlError("unterminated comment")
                self.i = e + 3
            elif self.starts("<?"):
                e = self.s.find("?>", self.i + 2)
                if e < 0:
                    raise XmlError("unterminated processing instruction")
                self.i = e + 2
            elif self.starts("<!"):
                e = self.s.find(">", self.i)
                if e < 0:
                    raise XmlError("unterminated declaration")
                self.i = e + 1
            else:
                return


def _parse_element(c: _Cursor, depth: int = 0) -> Node:
    if depth > 256:
        raise XmlError("nesting too deep")
    c.i += 1  # '<'
    b = c.i
    while not c.eof() and _is_name(c.s[c.i]):
        c.i += 1
    node = Node(c.s[b:c.i])
    if not node.tag:
        raise XmlError(f"empty tag name at offset {b}")
    while True:  # attributes; whitespace between them is optional
        c.skip_ws()
        if c.eof():
            raise XmlError(f"unterminated tag <{node.tag}")
        if c.starts("/>"):
            c.i += 2
            return node
        if c.s[c.i] == ">":
            c.i += 1
            break
        kb = c.i
        while not c.eof() and _is_name(c.s[c.i]):
            c.i += 1
        key = c.s[kb:c.i]
        if not key:
            raise XmlError(f"bad attribute in <{node.tag}> at offset {c.i}")
        c.skip_ws()
        val = ""
        if not c.eof() and c.s[c.i] == "=":
            c.i += 1
            c.skip_ws()
            if c.eof():
                raise XmlError("dangling '='")
            q = c.s[c.i]
            if q in "\"'":
                e = c.s.find(q, c.i + 1)
                if e < 0:
                    raise XmlError("unterminated attribute value")
                val = c.s[c.i + 1:e]
                c.i = e + 1
            else:
                vb = c.i
                while not c.eof() and not c.s[c.i].isspace() and c.s[c.i] != ">" and not c.starts("/>"):
                    c.i += 1
                val = c.s[vb:c.i]
        if key in node.attrs:
            raise XmlError(f"attribute '{key}' given twice on <{node.tag}>")
        node.attrs[key] = val
    while True:
        c.skip_misc()
        if c.eof():
            raise XmlError(f"missing </{node.tag}>")
        if c.starts("</"):
            e = c.s.find(">", c.i)
            if e < 0:
                raise XmlError("unterminated close tag")
            c.i = e + 1
            return node
        node.children.append(_parse_element(c, depth + 1))


def parse(text: str) -> Node:
    c = _Cursor(text)
    c.skip_misc()
    if c.eof():
        raise XmlError("no root element")
    return _parse_element(c)


def parse_file(path) -> Node:
    with open(path, "r") as f:
        return parse(f.read())


def _esc(v: str) -> str:
    return (str(v).replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;").replace('"', "&quot;"))


def dumps(node: Node, indent: int = 4, header: bool = True) -> str:
    out: List[str] = []
    if header:
        out.append('<?xml version="1.0" encoding="utf-8"?>')

    def rec(n: Node, d: int) -> None:
        pad = " " * (indent * d)
        attrs = "".join(f' {k}="{_esc(v)}"' for k, v in n.attrs.items())
        if n.children:
            out.append(f"{pad}<{n.tag}{attrs}>")
            for ch in n.children:
                rec(ch, d + 1)
            out.append(f"{pad}</{n.tag}>")
        else:
            out.append(f"{pad}<{n.tag}{attrs}/>")

    rec(node, 0)
    return "\n".join(out) + "\n"


def dump_file(node: Node, path, **kw) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write(dumps(node, **kw))
