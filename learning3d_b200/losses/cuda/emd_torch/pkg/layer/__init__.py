from .emd_loss_layer import EMDLoss, EMDFunction
