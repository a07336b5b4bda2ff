import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')
    # the shared library and the C oracle are built artefacts (git-ignored): build them when a fresh checkout
    # runs the tests before __graft_entry__.build() (nvcc / gcc cross-compile without a GPU)
    lib = os.path.join(ROOT, 'pasio_b200', 'libpasio_b200.so')
    ora = os.path.join(ROOT, 'oracle', '_build', 'libdp_oracle.so')
    if not (os.path.exists(lib) and os.path.exists(ora)):
        import __graft_entry__
        __graft_entry__.build()


_FIXTURE_CHECKS = {'bit_exact': 0, 'degraded': 0, 'labels_degraded': []}


def pytest_report_header(config):
    """Says up front whether comparisons against the reference-generated fixtures are bit-level on this host."""
    try:
        host = host_tables_sha1()
        fix = str(np.load(os.path.join(GOLDEN, 'exact.npz'), allow_pickle=False)['tables_sha1'])
    except Exception as e:          # noqa: BLE001
        return 'pasio_b200 fixtures: table hash unavailable (%s)' % e
    return ('pasio_b200 fixtures: host np.log/gammaln table sha1 %s, fixtures made with %s -> reference fixtures are '
            'compared %s' % (host[:12], fix[:12], 'BIT FOR BIT' if host == fix else
                             'score-only (1e-9); oracle-vs-CUDA comparisons on this host stay bit-exact'))


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    c = _FIXTURE_CHECKS
    line = ('reference-fixture comparisons: %d bit-exact, %d degraded to score-only (host tables differ from the '
            'fixtures\')' % (c['bit_exact'], c['degraded']))
    terminalreporter.write_line(line)
    try:
        out = os.path.join(ROOT, 'gpurun_out')
        os.makedirs(out, exist_ok=True)
        import json
        with open(os.path.join(out, 'fixture_parity.json'), 'a') as f:
            f.write(json.dumps({'host_tables_sha1': _HOST_SHA, 'bit_exact': c['bit_exact'], 'degraded': c['degraded'],
                                'degraded_labels': c['labels_degraded'][:20], 'markexpr': config.getoption('-m')}) + '\n')
    except OSError:
        pass


def host_tables_sha1():
    import scipy.special
    k = np.arange(1 << 20)
    h = hashlib.sha1()
    h.update(np.log(k + 1.0).tobytes())
    h.update(scipy.special.gammaln(k + 0).tobytes())
    return h.hexdigest()


_HOST_SHA = None


class Golden(object):
    """Fixture file made by oracle/make_golden.py from the unmodified reference."""

    def __init__(self, name):
        self.d = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
        global _HOST_SHA
        if _HOST_SHA is None:
            _HOST_SHA = host_tables_sha1()
        # bit-level comparison with the fixtures is meaningful only if this host's np.log /
        # gammaln produce the table bits the fixtures were generated with
        self.same_tables = ('tables_sha1' not in self.d.files) or (str(self.d['tables_sha1']) == _HOST_SHA)

    def __getitem__(self, key):
        return self.d[key]

    def bit_exact(self, label=''):
        """True when this comparison can be made bit for bit (host tables == fixture tables); every call is
        counted and the totals are printed at the end of the run, so a pass count never hides a degraded run."""
        if self.same_tables:
            _FIXTURE_CHECKS['bit_exact'] += 1
        else:
            _FIXTURE_CHECKS['degraded'] += 1
            _FIXTURE_CHECKS['labels_degraded'].append(label)
        return self.same_tables

    def has(self, key):
        return key in self.d.files

    def counts(self, name):
        from pasio_b200 import synth
        if self.has(name + '.counts'):
            return self.d[name + '.counts'].astype(np.int64)
        c = eval(str(self.d[name + '.gen']), {k: getattr(synth, k) for k in dir(synth)})
        assert hashlib.sha1(c.tobytes()).hexdigest() == str(self.d[name + '.counts_sha1'])
        return c

    def check_splits(self, got, want, score_got, score_want, what=''):
        """bit-exact on a host with the fixture's tables; otherwise score-only with 1e-9 relative"""
        assert abs(float(score_got) - float(score_want)) <= 1e-9 * abs(float(score_want)), what
        if self.bit_exact(what):
            assert np.array_equal(got, want), what
            assert float(score_got) == float(score_want), what


@pytest.fixture(scope='session')
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return get
