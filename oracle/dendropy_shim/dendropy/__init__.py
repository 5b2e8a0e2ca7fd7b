"""Minimal stand-in for the `dendropy` package -- TEST INFRASTRUCTURE ONLY.

The reference (torchtree) imports dendropy at module top
(torchtree/evolution/tree_model.py:9, io.py:6-9, cli/evolution.py:11) and uses
it for newick parsing, tree iteration and polytomy resolution only -- no
arithmetic.  dendropy is not installed in this image and cannot be installed
(no network), so this stand-in provides exactly the surface the reference
touches.  It exists so that `tests/golden/make_golden.py` can import the real
reference from /root/reference and so that CPU tests can run the reference
side by side with the plug-in.  It is never imported by the product package.
"""
from __future__ import annotations

import re

__version__ = "0.0-shim"


class Taxon:
    def __init__(self, label):
        self.label = label

    def __str__(self):
        # dendropy renders taxa quoted; callers strip the quotes
        # (tree_model.py:61, :228)
        return "'{}'".format(self.label)

    def __repr__(self):
        return "<Taxon {}>".format(self.label)


class TaxonNamespace:
    def __init__(self, labels=None):
        self._taxa = []
        self._by_label = {}
        if labels is not None:
            for label in labels:
                self.require_taxon(label)

    def require_taxon(self, label):
        if label not in self._by_label:
            taxon = Taxon(label)
            self._by_label[label] = taxon
            self._taxa.append(taxon)
        return self._by_label[label]

    def __len__(self):
        return len(self._taxa)

    def __iter__(self):
        return iter(self._taxa)

    def __getitem__(self, i):
        return self._taxa[i]

    def labels(self):
        return [t.label for t in self._taxa]


class _Annotations:
    def add_bound_attribute(self, name):
        pass


class Node:
    def __init__(self):
        self._children = []
        self.parent_node = None
        self.edge_length = None
        self.taxon = None
        self.label = None
        self.annotations = _Annotations()

    def is_leaf(self):
        return len(self._children) == 0

    def is_internal(self):
        return len(self._children) > 0

    def child_nodes(self):
        return list(self._children)

    def child_node_iter(self, filter_fn=None):
        for c in self._children:
            if filter_fn is None or filter_fn(c):
                yield c

    def num_child_nodes(self):
        return len(self._children)

    def add_child(self, node):
        node.parent_node = self
        self._children.append(node)
        return node

    def set_child_nodes(self, nodes):
        self._children = []
        for n in nodes:
            self.add_child(n)

    def postorder_iter(self, filter_fn=None):
        # iterative post-order (deep caterpillar trees exceed the recursion limit)
        stack = [(self, False)]
        while stack:
            node, expanded = stack.pop()
            if expanded or not node._children:
                if filter_fn is None or filter_fn(node):
                    yield node
            else:
                stack.append((node, True))
                for c in reversed(node._children):
                    stack.append((c, False))

    def preorder_iter(self, filter_fn=None):
        stack = [self]
        while stack:
            node = stack.pop()
            if filter_fn is None or filter_fn(node):
                yield node
            for c in reversed(node._children):
                stack.append(c)


_TOKEN = re.compile(r"\s*([(),;:]|\[[^\]]*\]|'(?:[^']|'')*'|[^(),;:\[\]\s']+)")


def _parse_newick(text, taxon_namespace):
    tokens = [t for t in _TOKEN.findall(text) if not t.startswith("[")]
    root = Node()
    current = root
    pos = 0
    last_closed = None
    expecting_length = False
    started = False
    for tok in tokens:
        pos += 1
        if tok == "(":
            if not started:
                started = True
                current = root
            child = Node()
            current.add_child(child)
            current = child
            last_closed = None
        elif tok == ",":
            parent = current.parent_node
            child = Node()
            parent.add_child(child)
            current = child
            last_closed = None
        elif tok == ")":
            current = current.parent_node
            last_closed = current
        elif tok == ":":
            expecting_length = True
        elif tok == ";":
            break
        else:
            if expecting_length:
                current.edge_length = float(tok)
                expecting_length = False
            else:
                label = tok
                if label.startswith("'"):
                    label = label[1:-1].replace("''", "'")
                if current.is_leaf() and last_closed is None:
                    current.taxon = taxon_namespace.require_taxon(label)
                else:
                    current.label = label
            if not started:
                started = True
    return root


class Tree:
    def __init__(self, seed_node, taxon_namespace):
        self.seed_node = seed_node
        self.taxon_namespace = taxon_namespace
        self.is_rooted = True

    @classmethod
    def get(
        cls,
        path=None,
        data=None,
        schema="newick",
        taxon_namespace=None,
        tree_offset=0,
        **kwargs
    ):
        if path is not None:
            with open(path) as fp:
                data = fp.read()
        if taxon_namespace is None:
            taxon_namespace = TaxonNamespace()
        if schema == "nexus":
            data = _newick_from_nexus(data, tree_offset)
        else:
            trees = [t for t in data.split(";") if t.strip()]
            data = trees[tree_offset] + ";"
        root = _parse_newick(data, taxon_namespace)
        root.edge_length = None
        return cls(root, taxon_namespace)

    # --- iteration -------------------------------------------------------
    def postorder_node_iter(self, filter_fn=None):
        return self.seed_node.postorder_iter(filter_fn)

    def postorder_internal_node_iter(self, filter_fn=None):
        for n in self.seed_node.postorder_iter(filter_fn):
            if not n.is_leaf():
                yield n

    def preorder_node_iter(self, filter_fn=None):
        return self.seed_node.preorder_iter(filter_fn)

    def leaf_node_iter(self, filter_fn=None):
        for n in self.seed_node.preorder_iter(filter_fn):
            if n.is_leaf():
                yield n

    def nodes(self):
        return list(self.preorder_node_iter())

    # --- editing ---------------------------------------------------------
    def resolve_polytomies(self, limit=2, update_bipartitions=False, rng=None):
        """Arbitrarily resolve polytomies with zero-length branches, in the
        deterministic (rng=None) way dendropy does: keep the first `limit-1`
        children... dendropy attaches a new node holding the *last* children;
        which zero-length edge is inserted does not change the likelihood."""
        polytomies = [
            n for n in self.postorder_node_iter() if len(n._children) > limit
        ]
        for node in polytomies:
            to_attach = node._children[limit:]
            for child in to_attach:
                node._children.remove(child)
            attachment_points = list(node._children)
            while to_attach:
                next_child = to_attach.pop()
                next_sib = attachment_points.pop()
                new_node = Node()
                new_node.edge_length = 0.0
                idx = node._children.index(next_sib)
                node._children[idx] = new_node
                new_node.parent_node = node
                new_node.add_child(next_sib)
                new_node.add_child(next_child)
                attachment_points.append(new_node)

    def as_string(self, schema="newick", **kwargs):
        def rec(n):
            if n.is_leaf():
                s = n.taxon.label
            else:
                s = "(" + ",".join(rec(c) for c in n._children) + ")"
            if n.edge_length is not None:
                s += ":{}".format(n.edge_length)
            return s

        return rec(self.seed_node) + ";"


def _newick_from_nexus(text, tree_offset=0):
    translate = {}
    m = re.search(r"translate(.*?);", text, re.IGNORECASE | re.DOTALL)
    if m:
        for entry in m.group(1).split(","):
            parts = entry.strip().split(None, 1)
            if len(parts) == 2:
                translate[parts[0]] = parts[1].strip().strip("'")
    trees = re.findall(
        r"^\s*tree\s+[^=]+=\s*(?:\[&[RU]\]\s*)?(.*?;)",
        text,
        re.IGNORECASE | re.MULTILINE | re.DOTALL,
    )
    newick = trees[tree_offset]
    if translate:
        newick = re.sub(
            r"(?<=[(,])\s*([^(),:;\[\]\s]+)",
            lambda mm: translate.get(mm.group(1), mm.group(1)),
            newick,
        )
    return newick


class _Sequence:
    def __init__(self, s):
        self._s = s

    def __str__(self):
        return self._s

    def symbols_as_string(self):
        return self._s

    def __len__(self):
        return len(self._s)


class DnaCharacterMatrix:
    def __init__(self, taxon_namespace):
        self.taxon_namespace = taxon_namespace
        self._seqs = {}

    @classmethod
    def get(cls, path=None, data=None, schema="fasta", taxon_namespace=None, **kwargs):
        if path is not None:
            with open(path) as fp:
                data = fp.read()
        if taxon_namespace is None:
            taxon_namespace = TaxonNamespace()
        self = cls(taxon_namespace)
        if schema != "fasta":
            raise NotImplementedError("dendropy shim: only fasta alignments")
        label = None
        chunks = []
        for line in data.splitlines():
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if label is not None:
                    self._seqs[label] = "".join(chunks)
                label = line[1:].strip()
                chunks = []
            else:
                chunks.append(line)
        if label is not None:
            self._seqs[label] = "".join(chunks)
        for label in self._seqs:
            taxon_namespace.require_taxon(label)
        return self

    def __len__(self):
        return len(self._seqs)

    def items(self):
        # taxon-namespace order
        for taxon in self.taxon_namespace:
            if taxon.label in self._seqs:
                yield taxon, _Sequence(self._seqs[taxon.label])

    def __getitem__(self, taxon):
        label = taxon.label if isinstance(taxon, Taxon) else taxon
        return _Sequence(self._seqs[label])
