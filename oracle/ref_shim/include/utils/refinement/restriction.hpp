// MOCK: multilevel restriction is outside what oracle/_ref compiles (uniform meshes only).
#pragma once
