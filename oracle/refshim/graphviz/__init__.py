"""Stand-in for graphviz (only imported by basecircuit.py for drawing; unused on the hot path)."""


class Graph:
    def __init__(self, *a, **k):
        raise ImportError("graphviz is not available")
