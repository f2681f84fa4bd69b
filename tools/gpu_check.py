"""Run every GPU test file in its own process (a CUDA fault poisons the context, so isolation keeps one bad
kernel from hiding the others), with a timeout each, logging to gpurun_out/."""
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / 'gpurun_out'
OUT.mkdir(exist_ok=True)
files = sys.argv[1:] or sorted(str(p.relative_to(ROOT)) for p in (ROOT / 'tests').glob('test_gpu_*.py'))
summary = []
for f in files:
    t = time.time()
    log = OUT / (Path(f).stem + '.log')
    try:
        r = subprocess.run([sys.executable, '-m', 'pytest', f, '-q', '-m', 'gpu', '-x' if '--x' in sys.argv else '-q',
                            '--tb=short', '-s'], cwd=ROOT, capture_output=True, text=True, timeout=900)
        log.write_text(r.stdout + '\n' + r.stderr)
        tail = [l for l in r.stdout.strip().splitlines() if l.strip()][-1:] or ['?']
        summary.append(f'{f}: rc={r.returncode} {time.time() - t:.0f}s :: {tail[0]}')
    except subprocess.TimeoutExpired as e:
        log.write_text((e.stdout or b'').decode(errors='ignore') + '\nTIMEOUT')
        summary.append(f'{f}: TIMEOUT')
    print(summary[-1], flush=True)
(OUT / 'gpu_check_summary.txt').write_text('\n'.join(summary) + '\n')
