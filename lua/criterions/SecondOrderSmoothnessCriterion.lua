-- model.lua:26 requires 'criterions.SecondOrderSmoothnessCriterion' by name (see OBGCCriterion.lua in this directory);
-- nn.SmoothnessCriterion and nn.SecondOrderSmoothnessCriterion are both defined by the Smoothness shim.
return require 'criterions.SmoothnessCriterion'
