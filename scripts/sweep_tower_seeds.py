import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import test_tower_gpu as t
for name in sorted(t.CASES):
    fails = []
    for seed in range(24):
        try:
            t.test_fused_tower_matches_float64_modules(name, seed)
        except AssertionError as e:
            fails.append((seed, ' | '.join(l.strip() for l in str(e).strip().splitlines() if 'Max' in l or 'fused' in l or 'ACTUAL' in l or 'DESIRED' in l or l.startswith('policy') or l.startswith('value') or l.startswith('return'))[:300]))
    print(name, 'failures', len(fails), fails[:8])
