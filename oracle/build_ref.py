"""Copy the unmodified reference into the git-ignored baseline/_ref/ so that the CPU-baseline legs of
bench.py can time the reference's own code on the GPU box (which has no /root/reference).
Nothing under baseline/_ref is tracked, imported by the product, or edited.  Usage: python -m oracle.build_ref
"""
import os
import shutil

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(src='/root/reference', dst=None):
    dst = dst or os.path.join(_REPO, 'baseline', '_ref')
    if not os.path.isdir(src):
        return False
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('.git', '__pycache__', '*.gif', 'res'))
    return True


if __name__ == '__main__':
    print('copied' if build() else 'reference not present; nothing done')
