import json, sys
d = json.load(open(sys.argv[1]))
B = d['chunk']; tot = sum(r['ms'] for r in d['rows']); peak = d['peaks']['bf16_sustained']
print('chunk', B, 'total us', round(tot * 1e3, 1), 'fps', round(B / tot * 1e3))
for r in d['rows']:
    tf = r['flops'] / r['ms'] / 1e9 if r['ms'] > 0 else 0
    fl = ('P' if r.get('fused_pool') else '-') + ('H' if r.get('halo') else '-') + ('F' if r.get('fused_head') else '-')
    print(f"{r['kind']:9s} {r['H']:4d}x{r['W']:<4d} {r['Cin']:5d}->{r['Cout']:<5d} bn={r['block_n']:<4d} {fl} {r['ms']*1e3:8.1f} us {tf:8.1f} TF/s  {100*tf/peak:5.1f}%  share {100*r['ms']/tot:4.1f}%")
