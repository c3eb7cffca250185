"""pip metadata; the CUDA library is built in-tree by `make -C numbskull_b200/csrc`
(or `python -c "import __graft_entry__ as g; g.build()"`), not by setuptools."""
from setuptools import setup, find_packages

exec(open('numbskull_b200/version.py').read())
setup(
    name='numbskull_b200',
    version=__version__,
    description='B200-native Gibbs sampling / weight learning behind the numbskull API',
    packages=find_packages(include=['numbskull_b200*', 'numbskull']),
    package_data={'numbskull_b200': ['libnumbskull_b200.so']},
    entry_points={'console_scripts': ['numbskull = numbskull_b200.numbskull:main']},
)
