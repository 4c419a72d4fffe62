"""Newick reader with the node numbering of NewickTree.jl#beluga's ``readnw`` (un-vendored
dependency of the reference, Manifest.toml:189-195): ids are pre-order, root = 1, children in
file order.  Call sites in the reference: src/model.jl:23,32."""


class NwNode:
    __slots__ = ("i", "name", "x", "p", "c")

    def __init__(self):
        self.i = 0
        self.name = ""
        self.x = float("nan")  # distance to parent
        self.p = None
        self.c = []


def _parse(s, pos):
    n = NwNode()
    while s[pos].isspace():
        pos += 1
    if s[pos] == "(":
        pos += 1
        while True:
            child, pos = _parse(s, pos)
            child.p = n
            n.c.append(child)
            while s[pos].isspace():
                pos += 1
            if s[pos] == ",":
                pos += 1
                continue
            if s[pos] == ")":
                pos += 1
                break
            raise ValueError("newick: expected ',' or ')' at %d" % pos)
    j = pos
    while j < len(s) and s[j] not in ":,);(":
        j += 1
    n.name = s[pos:j].strip()
    pos = j
    if pos < len(s) and s[pos] == ":":
        j = pos + 1
        while j < len(s) and s[j] not in ",);":
            j += 1
        n.x = float(s[pos + 1 : j])
        pos = j
    return n, pos


def readnw(s):
    """-> (t, l): root node and Dict id => leaf name."""
    s = s.strip()
    root, _ = _parse(s, 0)
    leaves = {}
    ctr = [0]

    def number(n):
        ctr[0] += 1
        n.i = ctr[0]
        if not n.c:
            leaves[n.i] = n.name
        for c in n.c:
            number(c)

    number(root)
    return root, leaves
