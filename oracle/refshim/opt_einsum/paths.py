def _chain(inputs, output, size_dict, memory_limit=None):
    """Contract, at every step, a pair of tensors that share an index (falls back to the first
    two): mirrors what a greedy finder does on a circuit network, in O(n^2)."""
    sets = [set(s) for s in inputs]
    path = []
    while len(sets) > 1:
        pick = None
        for i in range(len(sets)):
            for j in range(i + 1, len(sets)):
                if sets[i] & sets[j]:
                    pick = (i, j)
                    break
            if pick:
                break
        if pick is None:
            pick = (0, 1)
        i, j = pick
        new = (sets[i] | sets[j])
        # indices that appear only in these two (and not in the output) are summed away
        others = set().union(*[s for k, s in enumerate(sets) if k not in (i, j)]) | set(output)
        new = {x for x in new if x in others or not (x in sets[i] and x in sets[j])}
        new = {x for x in (sets[i] | sets[j]) if x in others} | {x for x in (sets[i] ^ sets[j])}
        sets = [s for k, s in enumerate(sets) if k not in (i, j)] + [new]
        path.append((i, j))
    return path


greedy = optimal = dynamic_programming = branch = auto = _chain
