"""CEL abstract syntax tree.

Host front-end data model. In the reference, policy conditions reach the engine
as cel-go ``CheckedExpr`` protos produced by ``StdEnv.Compile``
(internal/compile/conditions.go:63-89); the Go integration walks those protos
(with ``source_info.macro_calls`` to recover macro call sites). This module is
the equivalent tree for this repo, produced by :mod:`cerbos_b200.cel.parser`.

Node kinds mirror cel.dev/expr ``Expr``: Const, Ident, Select, Call, ListLit,
MapLit.  Macros (has/all/exists/exists_one/map/filter/two-variable
comprehensions/cel.bind/sortBy) are kept as high-level ``Macro`` nodes instead
of being expanded into fold comprehensions; their fold semantics (error
absorption of all/exists) are implemented by whoever consumes the tree.
"""
from __future__ import annotations


class UInt(int):
    """A CEL ``uint`` value (Python int tagged as unsigned)."""
    __slots__ = ()

    def __repr__(self):
        return f"{int(self)}u"


class Node:
    __slots__ = ()


class Const(Node):
    """Literal. value is: None, bool, int (CEL int), UInt, float, str, bytes."""
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value

    def __repr__(self):
        return f"Const({self.value!r})"


class Ident(Node):
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"Ident({self.name})"


class Select(Node):
    """operand.field ; test_only => has(operand.field)."""
    __slots__ = ("operand", "field", "test_only")

    def __init__(self, operand, field, test_only=False):
        self.operand = operand
        self.field = field
        self.test_only = test_only

    def __repr__(self):
        t = "has:" if self.test_only else ""
        return f"Select({t}{self.operand!r}.{self.field})"


class Call(Node):
    """Function call. Operators use cel-go's internal names (``_==_``, ``_&&_``,
    ``_?_:_``, ``_[_]``, ``@in``, ``!_``, ``-_`` ...). ``target`` is the receiver of
    a member call or None for a global call."""
    __slots__ = ("fn", "target", "args")

    def __init__(self, fn, target, args):
        self.fn = fn
        self.target = target
        self.args = list(args)

    def __repr__(self):
        if self.target is not None:
            return f"Call({self.target!r}.{self.fn}{self.args!r})"
        return f"Call({self.fn}{self.args!r})"


class ListLit(Node):
    __slots__ = ("elems",)

    def __init__(self, elems):
        self.elems = list(elems)

    def __repr__(self):
        return f"List{self.elems!r}"


class MapLit(Node):
    __slots__ = ("entries",)

    def __init__(self, entries):
        self.entries = list(entries)  # [(key_node, value_node)]

    def __repr__(self):
        return f"Map{self.entries!r}"


class Macro(Node):
    """High-level macro call.

    name: one of all, exists, exists_one, map, filter (1-var: vars=[x]);
          all, exists, existsOne, transformList, transformMap,
          transformMapEntry (2-var: vars=[i, v]); bind (vars=[name],
          target=init expr, args=[body]); sortBy (vars=[x], args=[key]).
    target: the range expression (or the bound init for ``bind``).
    args: [pred] | [transform] | [pred, transform]
    """
    __slots__ = ("name", "target", "vars", "args")

    def __init__(self, name, target, vars, args):
        self.name = name
        self.target = target
        self.vars = list(vars)
        self.args = list(args)

    def __repr__(self):
        return f"Macro({self.name} {self.target!r} {self.vars} {self.args!r})"


def walk(node):
    """Pre-order traversal of every node in the tree."""
    yield node
    if isinstance(node, Select):
        yield from walk(node.operand)
    elif isinstance(node, Call):
        if node.target is not None:
            yield from walk(node.target)
        for a in node.args:
            yield from walk(a)
    elif isinstance(node, ListLit):
        for e in node.elems:
            yield from walk(e)
    elif isinstance(node, MapLit):
        for k, v in node.entries:
            yield from walk(k)
            yield from walk(v)
    elif isinstance(node, Macro):
        yield from walk(node.target)
        for a in node.args:
            yield from walk(a)
