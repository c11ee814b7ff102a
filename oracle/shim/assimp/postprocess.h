#pragma once
#define aiProcessPreset_TargetRealtime_Quality 0
