def auto(nodes, output_edge_order=None, ignore_edge_order=False, **kws):
    raise NotImplementedError("tn.contractors.auto is off the golden path")


greedy = optimal = branch = auto
