import cProfile, pstats, sys, io
sys.path.insert(0, ".")
sys.argv = ["bench_configs.py", "3"]
import runpy
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path("scripts/bench_configs.py", run_name="__main__")
except SystemExit:
    pass
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(40)
