"""One-off wider differential fuzz: tests/test_gpu_tracker.py::test_tracker_fuzz_vs_oracle for seeds beyond the
ten of the test suite, and larger LSAP matrices of the contested / tied kinds.  python tools/fuzz_more.py [first] [count]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_tracker as T
import test_gpu_ops as O

first = int(sys.argv[1]) if len(sys.argv) > 1 else 10
count = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bad = []
for seed in range(first, first + count):
    try:
        T.test_tracker_fuzz_vs_oracle(seed)
    except AssertionError as e:
        bad.append((seed, str(e)[:200]))
print("tracker fuzz seeds %d..%d: %d failures %s" % (first, first + count - 1, len(bad), bad[:3]))
bad2 = []
for n in (5, 31, 33, 65, 97, 129, 180, 257, 400, 513, 640):
    try:
        O.test_lsap_known_first_step_shortcut_is_exact(n)
    except AssertionError as e:
        bad2.append((n, str(e)[:200]))
print("lsap shortcut sizes: %d failures %s" % (len(bad2), bad2[:3]))
sys.exit(1 if bad or bad2 else 0)
