import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_gpu_umma as T
for variant in (3, 4):
    for N, K in ((16, 128), (16, 16), (48, 64), (16, 32)):
        try:
            rel, err = T._run(N, K, variant=variant)
            print("variant", variant, "N", N, "K", K, "rel err %.3e" % rel, "timeout" if err else "")
        except Exception as e:
            print("variant", variant, N, K, "EXC", e)
