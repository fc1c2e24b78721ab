"""Result files in the reference's schema, so that runs produced with the CUDA estimators (and the
batched sampler front-end) can be read by the reference's R scripts (`r/helpers.R:1-108`).

Mirrors `parameter/mcmc/output.py:266-356` (`compile_results`, `save_to_file`) and
`helpers/file_system.py:43-77` (`write_to_json`): three gzip-compressed JSON files per simulation,

    <path>/<sim_name>/mcmc_output.json.gz   simulation_name, simulation_time, time_per_iteration and one
                                            (no_iters - no_burnin_iters) x dim table for each of params,
                                            params_prop, nat_gradient, accepted, state_trajectory that the
                                            first state of the history holds; NaN -> 0, +-inf -> +-1e7
    <path>/<sim_name>/data.json.gz          observations (and states when the model has none, as the
                                            reference writes it), simulation_name, simulation_time
    <path>/<sim_name>/settings.json.gz      every sampler setting as str(), sampler_name, simulation_*

plus `description.txt.gz` when a description is given and `benchmark.json.gz` when the states carry the
quasi-Newton benchmark errors (`output.py:212-263`).  The sampler object only has to offer `settings`,
`state_history` (dict or list of state dicts indexed by iteration), `time_per_iter`, `model` (`obs`,
`states`) and `name`.  Host-side Python: nothing here touches the device.
"""
import copy
import gzip
import json
import os
import time

import numpy as np

_TABLE_FIELDS = ('params', 'params_prop', 'nat_gradient', 'accepted', 'state_trajectory')
_BENCHMARK_FIELDS = ('error_bfgs_fro', 'error_ls_fro', 'error_sr1_fro')
_BIG = 10000000.0


def _clean(value):
    """NaN -> 0, +inf -> 1e7, -inf -> -1e7 (output.py:295-308), scalars and arrays alike."""
    arr = np.array(value, dtype=np.float64, copy=True).reshape(-1)
    arr[np.isnan(arr)] = 0.0
    arr[np.isposinf(arr)] = _BIG
    arr[np.isneginf(arr)] = -_BIG
    return arr


def _table(history, first, count, field, probe):
    """(count x dim) table of `field` over the iterations first .. first + count - 1; the width is
    taken from the state `probe` as the reference does."""
    width = 1 if np.isscalar(history[probe][field]) else len(history[probe][field])
    out = np.zeros((count, width))
    for i in range(count):
        out[i, :] = _clean(history[first + i][field])
    return out


def compile_results(mcmc, sim_name=None, now=None):
    """-> (mcmc_output dict, data dict, settings dict)"""
    no_iters = int(mcmc.settings['no_iters'])
    burn = int(mcmc.settings['no_burnin_iters'])
    keep = no_iters - burn
    stamp = now if now is not None else time.strftime("%c")
    history = mcmc.state_history

    out = {'simulation_name': sim_name, 'simulation_time': stamp, 'time_per_iteration': mcmc.time_per_iter}
    for field in _TABLE_FIELDS:
        if field in history[0]:
            out[field] = _table(history, burn, keep, field, 0)
    for attr in ('adapted_step_sizes', 'no_hessians_corrected'):
        if hasattr(mcmc, attr):
            out[attr] = getattr(mcmc, attr)

    data = {'observations': mcmc.model.obs}
    if mcmc.model.states is None:          # (sic) output.py:324: the states are written when there are none
        data['states'] = mcmc.model.states
    data['simulation_name'] = sim_name
    data['simulation_time'] = stamp

    settings = {key: str(val) for key, val in copy.deepcopy(mcmc.settings).items()}
    settings.update({'sampler_name': mcmc.name, 'simulation_name': sim_name, 'simulation_time': stamp})
    return out, data, settings


def compile_benchmark_results(mcmc, sim_name=None, now=None):
    no_iters = int(mcmc.settings['no_iters'])
    burn = int(mcmc.settings['no_burnin_iters'])
    history = mcmc.state_history
    last = no_iters - 1
    if 'error_sr1_fro' not in history[last]:
        return None
    stamp = now if now is not None else time.strftime("%c")
    out = {'simulation_name': sim_name, 'simulation_time': stamp, 'time_per_iteration': mcmc.time_per_iter}
    for field in _BENCHMARK_FIELDS:
        if field in history[last]:
            out[field] = _table(history, burn, no_iters - burn, field, last)
    return out


def write_to_json(data, output_path, sim_name, output_type, as_gzip=True):
    """helpers/file_system.py:43-77: arrays become lists, NaN / Infinity literals are allowed."""
    data = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in data.items()}
    file_name = os.path.join(output_path, sim_name, output_type)
    os.makedirs(os.path.dirname(file_name), exist_ok=True)
    if as_gzip:
        with gzip.GzipFile(file_name + '.gz', 'w') as fout:
            fout.write(json.dumps(data, allow_nan=True).encode('utf-8'))
    else:
        with open(file_name, 'w') as fout:
            json.dump(data, fout, ensure_ascii=False)
    return file_name + ('.gz' if as_gzip else '')


def save_to_file(mcmc, file_path, sim_name=None, sim_desc=None, now=None):
    """Writes the files listed in the module docstring; returns their paths."""
    out, data, settings = compile_results(mcmc, sim_name=sim_name, now=now)
    written = []
    if sim_desc:
        written.append(write_to_json({'description': sim_desc, 'time': settings['simulation_time']},
                                     file_path, sim_name, 'description.txt'))
    written.append(write_to_json(out, file_path, sim_name, 'mcmc_output.json'))
    written.append(write_to_json(data, file_path, sim_name, 'data.json'))
    written.append(write_to_json(settings, file_path, sim_name, 'settings.json'))
    bench = compile_benchmark_results(mcmc, sim_name=sim_name, now=now)
    if bench is not None:
        written.append(write_to_json(bench, file_path, sim_name, 'benchmark.json'))
    return written
