"""Import alias: `srba_stereo_slam_b200` -> the package directory `srba-stereo-slam_b200/`
(the hyphenated name the project layout prescribes is not a valid Python identifier)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "srba-stereo-slam_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
