import json
import sys
d = json.load(open(sys.argv[1]))
print('value', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2))
r = d['roofline']
print('  lstm_tc', round(r['mean_launch_ms'], 4), 'ms  share', round(r['share_of_step'], 4), ' TF/s', round(r['achieved'], 1))
for k, v in r.get('other_tc_kernels', {}).items():
    print('  %-10s %.4f ms  share %.4f  launches/step %d' % (k, v['mean_launch_ms'], v['share_of_step'],
                                                         v.get('launches_per_step', v.get('launches', 0))))
