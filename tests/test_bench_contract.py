"""CPU: the bench.py reference arm prints the contract line; the product package never
touches the oracle."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--qubits', '14', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
                'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['value'] > 0 and line['vs_baseline'] is None
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['value'] == line['value']
    assert 'workload' in line['config']


def test_product_never_imports_the_oracle():
    pattern = re.compile(r'^\s*(from|import)\s+(oracle|pauli_oracle|c_oracle|ref_shim)\b', re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'openfermion_b200')):
        for name in files:
            if name.endswith('.py'):
                text = open(os.path.join(dirpath, name)).read()
                assert not pattern.search(text), os.path.join(dirpath, name)
                assert 'oracle/' not in text.replace('oracle/make_golden.py', '').replace('oracle/)', ''), name
    entry = open(os.path.join(ROOT, '__graft_entry__.py')).read()
    assert 'import pauli_oracle' in entry  # smoke() is allowed to check against it
