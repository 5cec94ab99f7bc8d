import json, sys, collections
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/profile_step.json'))
L = d['shapes']
b = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (name, tc, M, Cin, Cout, T, splits), (cnt, ms, fl) in L:
    key = (name, tc, 'M<=4096' if M <= 4096 else ('M<=32768' if M <= 32768 else 'big'))
    ideal = max(fl / 1390e12, cnt * (M * Cin * 2 + M * Cout * 4) / 6.5e12) * 1e3
    b[key][0] += cnt; b[key][1] += ms; b[key][2] += fl; b[key][3] += ideal
for k, v in sorted(b.items(), key=lambda kv: -kv[1][1]):
    print(k, 'n=%d %.1f ms  %.2f TFLOP  %.0f TF/s  ideal %.2f ms' % (v[0], v[1], v[2] / 1e12, v[2] / v[1] / 1e9, v[3]))
print("--- small-M convs by time")
rows = [(k, v) for k, v in L if k[0] == 'conv_gemm' and k[2] <= 4096]
for k, v in sorted(rows, key=lambda kv: -kv[1][1])[:25]:
    print(k, 'n=%d %.2f ms  %.1f us/launch %.0f TF/s' % (v[0], v[1], 1e3 * v[1] / v[0], v[2] / v[1] / 1e9))
