"""Differentiable (create_graph) evaluation for training - SURVEY.md section 8a row T.

Not part of the round-1 inference path; NewtonNet.forward routes here when a derivative head has
create_graph=True (model.train()).
"""


def differentiable_forward(model, z, pos, cell, batch):
    raise NotImplementedError(
        'training-mode (create_graph=True) evaluation is not implemented yet in newtonnet_b200; call '
        'model.eval() for energy / force / stress inference')
