"""Print the headline fields of bench.py JSON lines read from stdin (flags anything else on stdout)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ''
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print(tag, 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'],
              'host_issue %.4f' % d.get('host_issue_ms_per_step', -1), 'single %s' % d.get('single_stream'),
              'e2e %.4g (%.3f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']),
              'roof %s %.3f' % (d['roofline']['kernel'], d['roofline']['frac']), 'launches', d.get('gpu_launches'))
    elif line.strip():
        print('STDOUT NOISE:', line[:80].rstrip())
