-- model.lua:28 requires 'criterions.OcclusionPriorCriterion' by name (see OBGCCriterion.lua in this directory);
-- nn.ConstVelCriterion and nn.OcclusionPriorCriterion are both defined by the ConstVel shim.
return require 'criterions.ConstVelCriterion'
