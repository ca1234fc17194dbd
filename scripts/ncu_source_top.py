#!/usr/bin/env python
"""Top SASS lines by stall samples + instruction totals from an `ncu --page source --csv` export."""
import csv
import sys


def main(path, rows_processed=None, top=22):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    tot_inst = sum(int(r[ci['Instructions Executed']]) for r in body)
    tot_samp = sum(int(r[ci['# Samples']]) for r in body)
    print(f'SASS lines {len(body)}  warp-instructions {tot_inst}  samples {tot_samp}')
    if rows_processed:
        print(f'  = {tot_inst * 32 / rows_processed:.1f} thread-instructions per row')
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(int(r[ci[h]]) for r in body) for h in stall_cols}
    print('  stalls: ' + ', '.join(f'{k[6:]}={v * 100 // max(1, tot_samp)}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    for r in sorted(body, key=lambda r: -int(r[ci['# Samples']]))[:top]:
        big = max(stall_cols, key=lambda h: int(r[ci[h]]))
        print(f"{int(r[ci['# Samples']]):7d} {int(r[ci['Instructions Executed']]):10d}  {r[ci['Source']].strip()[:70]:70s} {big[6:]}")


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
