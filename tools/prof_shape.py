import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conv_sweep
name, tcv = sys.argv[1], int(sys.argv[2])
B, T, ci, co, k, d = conv_sweep.shapes[name]
print(conv_sweep.time_conv(B, T, ci, co, k, d, tcv, iters=2, res=(name == "res2")))
