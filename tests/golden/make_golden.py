"""Extract the golden values the reference stores in its own notebooks.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):  python tests/golden/make_golden.py
Writes tests/golden/reference_golden.json and tests/golden/mannheim_quad.npz.

The reference cannot be executed here (jax is not installed and there is no
network), so the stored cell outputs of Test/*.ipynb are the reference results
that pin the oracle.
"""
import json
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def cells(nb):
    with open(os.path.join(REF, nb)) as f:
        return json.load(f)['cells']


def out_text(cell):
    txt = []
    for o in cell.get('outputs', []):
        if 'text' in o:
            txt.append(''.join(o['text']))
        elif 'data' in o and 'text/plain' in o['data']:
            txt.append(''.join(o['data']['text/plain']))
    return '\n'.join(txt)


FLOAT = r'[-+]?\d+\.\d+(?:[eE][-+]?\d+)?'


def floats(s):
    return [float(x) for x in re.findall(FLOAT, s)]


gold = {}

# G1/G2: barrel arch FEA -- Test/shells_fea_validation.ipynb cells 4-7
c = cells('Test/shells_fea_validation.ipynb')
gold['shell_arch_min_uz'] = {
    'dense': floats(out_text(c[4]))[0], 'scipy': floats(out_text(c[5]))[0],
    'jax_sparse': floats(out_text(c[6]))[0],
    'source': 'Test/shells_fea_validation.ipynb cells 4,5,6'}
gold['shell_arch_strain_energy'] = {
    'value': floats(out_text(c[7]))[0], 'source': 'Test/shells_fea_validation.ipynb cell 7'}

# G3: barrel arch dC/dz at node 201 -- Test/shells_ad_validation.ipynb cell 13
c = cells('Test/shells_ad_validation.ipynb')
v = floats(out_text(c[13]))
gold['shell_arch_grad_node201'] = {
    'dense': v[0], 'jax_sparse': v[1], 'scipy': v[2], 'design_i': 201,
    'source': 'Test/shells_ad_validation.ipynb cell 13'}

# G4/G5: beam arch FEA -- Test/beamcols_fea_validation.ipynb cells 4, 5
c = cells('Test/beamcols_fea_validation.ipynb')
v4, v5 = floats(out_text(c[4])), floats(out_text(c[5]))
gold['beam_arch'] = {
    'dense_min_uz': v4[0], 'dense_strain_energy': v4[1],
    'scipy_min_uz': v5[0], 'scipy_strain_energy': v5[1],
    'source': 'Test/beamcols_fea_validation.ipynb cells 4,5'}

# G6: beam arch dC/dz at node 49 -- Test/beamcols_ad_validation.ipynb cell 12
c = cells('Test/beamcols_ad_validation.ipynb')
v = floats(out_text(c[12]))
gold['beam_arch_grad_node49'] = {
    'dense': v[0], 'jax_sparse': v[1], 'scipy': v[2], 'design_i': 49,
    'source': 'Test/beamcols_ad_validation.ipynb cell 12'}

# G7: Frames f(n,100) min u_z -- Test/Frames_speed.ipynb; each run prints
#   <ms>\nEach span has n elements\n<min uz>\nDOF ...
# the elapsed-ms line is printed before or after the 'Each span' line depending on the cell
SPAN = r'Each span has (\d+) elements\n(?:' + FLOAT + r'\n)?(-' + r'\d+\.\d+(?:[eE][-+]?\d+)?' + r')\nDOF'
frames = {}
for cell in cells('Test/Frames_speed.ipynb'):
    if cell['cell_type'] != 'code':
        continue
    txt = out_text(cell)
    for m in re.finditer(SPAN, txt):
        frames[int(m.group(1))] = float(m.group(2))  # later (post-JIT) run wins; same value
gold['frames_min_uz'] = {'by_n': {str(k): frames[k] for k in sorted(frames)},
                         'source': 'Test/Frames_speed.ipynb (printed per f(n,100) run)'}

# G8: Frames gradient minima -- Test/Frames_speed-Sensitivity.ipynb
sens = {}
for cell in cells('Test/Frames_speed-Sensitivity.ipynb'):
    if cell['cell_type'] != 'code':
        continue
    txt = out_text(cell)
    for m in re.finditer(SPAN, txt):
        sens.setdefault(int(m.group(1)), float(m.group(2)))
gold['frames_min_grad'] = {'by_n': {str(k): sens[k] for k in sorted(sens)},
                           'source': 'Test/Frames_speed-Sensitivity.ipynb (printed per run)'}

with open(os.path.join(HERE, 'reference_golden.json'), 'w') as f:
    json.dump(gold, f, indent=1)
print(json.dumps(gold, indent=1))

# Mannheim_Quad mesh data (Examples/Data/Mannheim_Quad/*.csv, read as in
# Examples/Shells_Mannheim_Multihalle_Shape.ipynb cell 4: header=None)
d = os.path.join(REF, 'Examples/Data/Mannheim_Quad')
ld = lambda n: np.loadtxt(os.path.join(d, n), ndmin=1)
raw = ld('cnct.csv').astype(np.int32)
n_ele = (raw.shape[0] + 1) // 4
np.savez_compressed(os.path.join(HERE, 'mannheim_quad.npz'),
                    x=ld('crd_x.csv'), y=ld('crd_y.csv'), z=ld('crd_z.csv'),
                    cnct=raw[:n_ele * 4].reshape(n_ele, 4),
                    bc_nodes=ld('bc_node.csv').astype(np.int32))
print('mannheim:', n_ele, 'quads', ld('crd_x.csv').shape[0], 'nodes')
