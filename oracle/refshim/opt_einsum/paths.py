def _chain(inputs, output, size_dict, memory_limit=None):
    """Greedy pairwise path: at every step contract the pair of tensors sharing an index whose
    result is smallest (ties: first found); falls back to the first two when nothing is shared.
    On a circuit network this keeps absorbing gates into the state instead of multiplying gates
    with each other.  O(n^3) in the number of tensors, fine for test-sized circuits."""
    sets = [set(s) for s in inputs]
    out = set(output)
    path = []
    while len(sets) > 1:
        # how many tensors (plus the output) hold each index
        count = {}
        for s in sets:
            for x in s:
                count[x] = count.get(x, 0) + 1
        best = None
        for i in range(len(sets)):
            for j in range(i + 1, len(sets)):
                shared = sets[i] & sets[j]
                if not shared:
                    continue
                summed = {x for x in shared if count[x] == 2 and x not in out}
                size = len((sets[i] | sets[j]) - summed)
                if best is None or size < best[0]:
                    best = (size, i, j, summed)
        if best is None:
            i, j, summed = 0, 1, set()
        else:
            _, i, j, summed = best
        new = (sets[i] | sets[j]) - summed
        sets = [s for k, s in enumerate(sets) if k not in (i, j)] + [new]
        path.append((i, j))
    return path


greedy = optimal = dynamic_programming = branch = auto = _chain
