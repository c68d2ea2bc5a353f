-- model.lua:22 requires 'criterions.OBGCCriterion' by name: without this file the reference's own Lua implementation
-- would be found further down package.path and redefine nn.OBGCCriterion.  Both classes live in the OBCC shim.
return require 'criterions.OBCCriterion'
