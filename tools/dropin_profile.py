"""Developer tool: cProfile of the reference-signature batched path (tools/dropin_bench.py)."""
import cProfile, pstats, io, os, sys
sys.argv = [sys.argv[0]]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
pr = cProfile.Profile()
pr.enable()
import dropin_bench  # noqa: F401  (runs main())
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
