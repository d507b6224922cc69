"""Instruction count per source line for one kernel of libprecond_b200.so (needs -lineinfo).

usage: python scripts/sass_lines.py <kernel-name-substring> [top]
"""
import collections, os, re, subprocess, sys, tempfile

def main():
  pat = sys.argv[1]
  top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
  so = os.path.join(os.path.dirname(__file__), "..", "precondition_b200", "libprecond_b200.so")
  with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in sorted(os.listdir(d)):
      txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True,
                           text=True).stdout
      cur, line = None, None
      cnt = collections.defaultdict(collections.Counter)
      for l in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
        if m:
          cur = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
          line = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
          cnt[cur][line] += 1
      for k, c in cnt.items():
        if pat not in k:
          continue
        print(k[:70], sum(c.values()), "instructions")
        for (fn, ln), n in sorted(c.items(), key=lambda x: -x[1])[:top]:
          print(f"   {fn}:{ln}  {n}")

main()
