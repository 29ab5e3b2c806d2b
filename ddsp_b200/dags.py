"""DAG executor with the semantics of `ddsp/dags.py` (dags.py:39-195).

Plain Python host glue (no keras): nodes are (module_or_name, [input keys],
[optional output keys]); outputs start as {'inputs': inputs, **inputs}; the last
node is aliased as 'out'.
"""
import logging
from typing import Dict, Sequence, Text, Tuple, Any

from ddsp_b200 import core

TensorDict = Dict[Text, Any]
Node = Tuple[Any, Sequence[Text], Sequence[Text]]
DAG = Sequence[Node]

# Helper functions (dags.py:39-54) --------------------------------------------
filter_by_value = lambda d, cond: dict(filter(lambda e: cond(e[1]), d.items()))
# Duck typing: anything callable with a `name` is a module here (the reference
# checks isinstance(v, tf.Module), dags.py:40).
is_module = lambda v: callable(v) and hasattr(v, 'name') and not isinstance(v, str)
is_loss = lambda v: hasattr(v, 'get_losses_dict')
is_processor = lambda v: hasattr(v, 'get_signal') and hasattr(v, 'get_controls')


def split_keras_kwargs(kwargs):
  """dags.py:47-53 - strip keras-only kwargs (kept for call compatibility)."""
  keras_kwargs = {}
  for key in ['training', 'mask', 'name']:
    if kwargs.get(key) is not None:
      keras_kwargs[key] = kwargs.pop(key)
  return keras_kwargs, kwargs


class DAGLayer:
  """String modules together (dags.py:57-195)."""

  def __init__(self, dag: DAG, **kwarg_modules):
    keras_kwargs, kwarg_modules = split_keras_kwargs(kwarg_modules)
    self.name = keras_kwargs.get('name', 'dag_layer')
    modules = filter_by_value(kwarg_modules, is_module)
    dag, dag_modules = self.format_dag(dag)
    self.dag = dag
    modules.update(dag_modules)
    self.module_names = list(modules.keys())
    for module_name, module in modules.items():
      setattr(self, module_name, module)

  @property
  def modules(self):
    return [getattr(self, name) for name in self.module_names]

  @staticmethod
  def format_dag(dag):
    """dags.py:108-127 - replace module instances by their names."""
    modules = {}
    dag = list(dag)
    for i, node in enumerate(dag):
      node = list(node)
      module = node[0]
      if is_module(module):
        modules[module.name] = module
        node[0] = module.name
      dag[i] = node
    return dag, modules

  def __call__(self, inputs: TensorDict, **kwargs):
    return self.call(inputs, **kwargs)

  def call(self, inputs: TensorDict, **kwargs):
    return self.run_dag(inputs, **kwargs)

  def run_dag(self, inputs: TensorDict, verbose: bool = False,
              **kwargs) -> TensorDict:
    """dags.py:134-195."""
    outputs = {'inputs': inputs}
    outputs.update(inputs)
    module_outputs = None
    for node in self.dag:
      module_key, input_keys = node[0], node[1]
      module = getattr(self, module_key)
      output_keys = node[2] if len(node) > 2 else None
      node_inputs = [core.nested_lookup(key, outputs) for key in input_keys]
      if verbose:
        logging.info('Input to Module: %s\nKeys: %s\nIn: %s\n', module_key,
                     input_keys, [list(getattr(x, 'shape', [])) for x in node_inputs])
      if is_processor(module):
        module_outputs = module(*node_inputs, return_outputs_dict=True, **kwargs)
      elif is_loss(module):
        module_outputs = module.get_losses_dict(*node_inputs, **kwargs)
      else:
        module_outputs = module(*node_inputs, **kwargs)
      if not isinstance(module_outputs, dict):
        module_outputs = core.to_dict(module_outputs, output_keys)
      if verbose:
        logging.info('Output from Module: %s\nOut keys: %s\n', module_key,
                     core.nested_keys(module_outputs))
      outputs[module_key] = module_outputs
    outputs['out'] = module_outputs
    return outputs
