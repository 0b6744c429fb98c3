"""Config C3 alone (A/B runs): time nans_check_collision_device over N random pairs, print ms and pairs/s; with a second
argument also check that many hit flags against the reference binary.  usage: python tools/c3_time.py [pairs] [check]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":          # the checker's worker processes re-import this file (spawn)
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
    chk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    r = bench.run_c3(n, 0, chk, reps=3)
    print(json.dumps({k: r[k] for k in ("pairs", "ms", "pairs_per_s", "hit_rate", "flags_bit_exact", "mismatches") if k in r}))
